"""Time ndconv_conv_fft_sharded_device on the BASELINE workload c5 (2-D f32 32768^2, k = 63^2, Full, Reflect) with the array ALREADY
device-resident and row-partitioned over all visible GPUs (one process, one processor handle and host thread per GPU): ghost-row
halo exchange over NVLink + the single-GPU pipeline in place on every shard.  Prints one JSON line (append to profiles/)."""
import importlib, json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
pkg = importlib.import_module("ndarray-conv_b200")

n0 = n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
reps = 10
k = np.random.default_rng(2005).random((63, 63), dtype=np.float32)
cm, pm = pkg.ConvMode.Full, pkg.PaddingMode.Reflect
rows = [n0 // ndev + (1 if g < n0 % ndev else 0) for g in range(ndev)]
procs = [pkg.get_fft_processor(d) for d in range(ndev)]
pl = [pkg.shard_plan((n0, n1), np.float32, k, cm, pm, rows, g) for g in range(ndev)]
bufs, outs, shards, streams = [], [], [], []
for g in range(ndev):
    dev = torch.device("cuda", g)
    hf, hb = pl[g]["halo_front"], pl[g]["halo_back"]
    gen = torch.Generator(device=dev); gen.manual_seed(1005 + g)
    buf = torch.rand((hf + rows[g] + hb, n1), dtype=torch.float32, device=dev, generator=gen)
    o = torch.empty((pl[g]["out_end"] - pl[g]["out_begin"], n1 + 62), dtype=torch.float32, device=dev)
    bufs.append(buf); outs.append(o)
    shards.append(dict(data=buf.data_ptr() + hf * n1 * 4, rows=rows[g], halo_front=hf, halo_back=hb, out=o.data_ptr()))
    st = torch.cuda.Stream(dev); streams.append(st); procs[g].set_stream(st.cuda_stream)
for d in range(ndev):
    torch.cuda.synchronize(d)

def call():
    pkg.conv_fft_sharded_device(procs, (n0, n1), np.float32, k, cm, pm, shards)

for _ in range(3):
    call()
for p in procs:
    p.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ndev)]
t0 = time.perf_counter()
for g in range(ndev):
    with torch.cuda.device(g):
        ev[g][0].record(streams[g])
for _ in range(reps):
    call()
for g in range(ndev):
    with torch.cuda.device(g):
        ev[g][1].record(streams[g])
for p in procs:
    p.synchronize()
wall = (time.perf_counter() - t0) / reps
dev_ms = max(ev[g][0].elapsed_time(ev[g][1]) for g in range(ndev)) / reps
total_out = (n0 + 62) * (n1 + 62)
halo_bytes = sum((p["halo_front"] + p["halo_back"]) * n1 * 4 for p in pl)
# the ghost rows really came from the neighbours: shard g's front ghost rows equal the last owned rows of shard g - 1
ok = True
for g in range(1, ndev):
    hf = pl[g]["halo_front"]
    if hf:
        a = bufs[g][:hf].cpu()
        hfp = pl[g - 1]["halo_front"]
        b = bufs[g - 1][hfp + rows[g - 1] - hf:hfp + rows[g - 1]].cpu()
        ok = ok and bool(torch.equal(a, b))
print(json.dumps({"what": "ndconv_conv_fft_sharded_device, c5-shaped, device-resident shards", "shape": [n0, n1], "n_gpus": ndev, "ms_per_call_device_max": dev_ms,
                  "ms_per_call_wall": wall * 1e3, "Gsamples_per_s": total_out / (dev_ms * 1e-3) / 1e9, "halo_bytes_per_call": halo_bytes,
                  "halo_rows": [(p["halo_front"], p["halo_back"]) for p in pl], "ghost_rows_equal_neighbours": ok, "reps": reps}))
for p in procs:
    p.close()
