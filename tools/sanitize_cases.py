"""a few small end-to-end calls covering every kernel family -- run under compute-sanitizer"""
import importlib, sys, numpy as np
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
rng = np.random.default_rng(0)
proc = pkg.get_fft_processor(0)
B = pkg.BorderType
# fast path 2-D (256-row tiles and 1024-row tiles), generic 1-D / 3-D / complex / f64, direct tile (TMA + patch) and generic direct (rank 4)
x = rng.random((300, 1400), dtype=np.float32); k = rng.random((5, 7), dtype=np.float32)
pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, proc)
x = rng.random((1100, 2100), dtype=np.float32)
pkg.conv_fft_with_processor(x, pkg.with_dilation(k, 2), pkg.ConvMode.Same, pkg.PaddingMode.Custom([B.Circular, B.Const(1.0)]), proc)
pkg.conv_fft_with_processor(rng.random(5000, dtype=np.float32), rng.random(31, dtype=np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Zeros, proc)
pkg.conv_fft_with_processor(rng.random((6, 20, 30)), rng.random((2, 3, 4)), pkg.ConvMode.Full, pkg.PaddingMode.Replicate, proc)
xc = (rng.random((20, 40)) + 1j * rng.random((20, 40))).astype(np.complex64)
pkg.conv_fft_with_processor(xc, xc[:3, :4].copy(), pkg.ConvMode.Same, pkg.PaddingMode.Circular, proc)
xi = rng.integers(-9, 9, size=(16, 40, 48), dtype=np.int32); ki = rng.integers(-3, 3, size=(3, 5, 5), dtype=np.int32)
pkg.conv(xi, ki, pkg.ConvMode.Custom([1, 2, 2], [2, 2, 2]), pkg.PaddingMode.Replicate, processor=proc)
pkg.conv(xi, ki, pkg.ConvMode.Same, pkg.PaddingMode.Zeros, processor=proc)
pkg.conv(xi.astype(np.int8), ki.astype(np.int8), pkg.ConvMode.Full, pkg.PaddingMode.Reflect, processor=proc)
x4 = rng.integers(-9, 9, size=(3, 4, 5, 6), dtype=np.int64)
pkg.conv(x4, x4[:2, :2, :2, :2].copy(), pkg.ConvMode.Same, pkg.PaddingMode.Circular, processor=proc)
# round 1b: rank-1 fused kernel (256- and 2048-sample tiles, strided), Complex<f32> fast path (rank 2 and 3), axis-0 split, 512-row column tiles,
# direct tile kernel without TMA (gathered window) and with a Const border patch
pkg.conv_fft_with_processor(rng.random(700, dtype=np.float32), rng.random(5, dtype=np.float32), pkg.ConvMode.Full, pkg.PaddingMode.Reflect, proc)
pkg.conv_fft_with_processor(rng.random(70000, dtype=np.float32), rng.random(63, dtype=np.float32), pkg.ConvMode.Custom([9], [3]), pkg.PaddingMode.Const(2.0), proc)
xc2 = (rng.random((300, 700)) + 1j * rng.random((300, 700))).astype(np.complex64)
pkg.conv_fft_with_processor(xc2, xc2[:5, :9].copy(), pkg.ConvMode.Full, pkg.PaddingMode.Custom([B.Reflect, B.Const(1 - 2j)]), proc)
xc3 = (rng.random((40, 90, 130)) + 1j * rng.random((40, 90, 130))).astype(np.complex64)
pkg.conv_fft_with_processor(xc3, xc3[:3, :4, :5].copy(), pkg.ConvMode.Same, pkg.PaddingMode.Replicate, proc)
pkg.conv_fft_with_processor(rng.random((1100, 4000), dtype=np.float32), k, pkg.ConvMode.Custom([2, 4], [3, 2]), pkg.PaddingMode.Replicate, proc)
pkg.conv_fft_with_processor(rng.random((700, 1500), dtype=np.float32), k, pkg.ConvMode.Same, pkg.PaddingMode.Zeros, proc)
pkg.conv(xi.astype(np.int16), ki.astype(np.int16), pkg.ConvMode.Custom([1, 2, 2], [2, 1, 2]), pkg.PaddingMode.Const(3), processor=proc)
pkg.conv(rng.random((16, 40, 48), dtype=np.float32), rng.random((3, 5, 5), dtype=np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Const(0.5), processor=proc)
s = proc.forward(rng.random((12, 40), dtype=np.float32)); proc.backward(s)
# round 1d: global-memory Stockham passes of the public processors (long strided axis, odd real axis, prime factor 13, complex,
# f64), the sharded host call (two handles on this device; large enough to take the pipelined host path)
for shp, dt in (((1100, 6), np.float32), ((5, 26, 9), np.float64), ((2, 9000), np.complex64), ((97,), np.float32)):
    xx = rng.random(shp).astype(dt)
    s = proc.forward(xx); proc.backward(s)
proc2 = pkg.get_fft_processor(0)
pkg.conv_fft_sharded(rng.random((3000, 4200), dtype=np.float32), k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, [proc, proc2])
proc2.close()
# round 2: TMA-fed 1024-row column pass (modes FWD / INV / FMI, row pitch 1032 and others), same-shape batch fold, i128 direct conv,
# a kernel longer than one FFT tile (cut into segments), Bluestein axes, device-resident shards with ghost rows (two handles)
pkg.conv_fft_with_processor(rng.random((1100, 2100), dtype=np.float32), k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, proc)
pkg.conv_fft_with_processor(rng.random((3, 1000, 300), dtype=np.float32), rng.random((2, 3, 5), dtype=np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Reflect, proc)
pkg.conv_fft_with_processor(rng.random((1000, 20, 300), dtype=np.float32), rng.random((3, 3, 3), dtype=np.float32), pkg.ConvMode.Full, pkg.PaddingMode.Replicate, proc)
pkg.conv_fft_with_processor(rng.random((5, 200, 700), dtype=np.float32), rng.random((1, 5, 7), dtype=np.float32), pkg.ConvMode.Same,
                            pkg.PaddingMode.Custom([B.Zeros, B.Reflect, B.Circular]), proc)
pkg.conv(pkg.int128_array(rng.integers(-9, 9, size=(6, 20, 24)).astype(object) * (1 << 70)), pkg.int128_array(rng.integers(-3, 3, size=(3, 3, 3)).astype(object)),
         pkg.ConvMode.Same, pkg.PaddingMode.Replicate, processor=proc)
pkg.conv_fft_with_processor(rng.random((2600, 40), dtype=np.float32), rng.random((1500, 3), dtype=np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Replicate, proc)
for shp, dt in (((2, 1009), np.float32), ((211, 4), np.complex64)):
    xx = rng.random(shp).astype(dt)
    s = proc.forward(xx); proc.backward(s)
import torch
procs = [pkg.get_fft_processor(0), pkg.get_fft_processor(0)]
kk = rng.random((9, 5), dtype=np.float32)
rows = [700, 800]
pl = [pkg.shard_plan((1500, 1300), np.float32, kk, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, rows, g) for g in range(2)]
bufs, outs, shards = [], [], []
for g in range(2):
    hf, hb = pl[g]["halo_front"], pl[g]["halo_back"]
    buf = torch.rand((hf + rows[g] + hb, 1300), dtype=torch.float32, device="cuda")
    o = torch.empty((pl[g]["out_end"] - pl[g]["out_begin"], 1304), dtype=torch.float32, device="cuda")
    bufs.append(buf); outs.append(o)
    shards.append(dict(data=buf.data_ptr() + hf * 1300 * 4, rows=rows[g], halo_front=hf, halo_back=hb, out=o.data_ptr()))
torch.cuda.synchronize()
pkg.conv_fft_sharded_device(procs, (1500, 1300), np.float32, kk, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, shards)
for q in procs:
    q.synchronize(); q.close()
# round 2b: Tensor-Memory-resident column pass (>= 10 tiles of 1024 rows: bundles, chunk hand-over between the two groups, parts through the
# scratch buffers; one geometry with fewer chunks than CTAs and one with a single bundle), f64 fast path (rank 2 with edge tiles, rank 3,
# strided output), register-blocked direct conv (run with NDCONV_BLOCKED_MIN_OUT=0: strides / dilations 1 and 2, 4- and 8-byte elements)
pkg.conv_fft_with_processor(rng.random((10200, 520), dtype=np.float32), rng.random((3, 9), dtype=np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Const(0.5), proc)
pkg.conv_fft_with_processor(rng.random((12000, 1000), dtype=np.float32), rng.random((5, 5), dtype=np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Replicate, proc)
pkg.conv_fft_with_processor(rng.random((700, 1500)), rng.random((5, 9)), pkg.ConvMode.Full, pkg.PaddingMode.Reflect, proc)
pkg.conv_fft_with_processor(rng.random((20, 70, 300)), rng.random((3, 4, 5)), pkg.ConvMode.Custom([1, 0, 5], [2, 1, 3]), pkg.PaddingMode.Custom([B.Reflect, B.Circular, B.Const(0.5)]), proc)
pkg.conv_fft_with_processor(rng.random((140, 150)), rng.random((3, 3)), pkg.ConvMode.Same, pkg.PaddingMode.Reflect, proc)
xb = rng.integers(-9, 9, size=(5, 40, 200), dtype=np.int32)
pkg.conv(xb, ki, pkg.ConvMode.Custom([1, 2, 2], [2, 2, 2]), pkg.PaddingMode.Replicate, processor=proc)
pkg.conv(xb.astype(np.int64), pkg.with_dilation(ki[:, :3, :4].astype(np.int64), 2), pkg.ConvMode.Same, pkg.PaddingMode.Circular, processor=proc)
pkg.conv(rng.random((64, 260), dtype=np.float32), rng.random((7, 7), dtype=np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Reflect, processor=proc)
pkg.conv(rng.random((20, 150)), rng.random((3, 4)), pkg.ConvMode.Explicit([[1, 1], [3, 2]], [1, 2]), pkg.PaddingMode.Reflect, processor=proc)
proc.close()
print("SANITIZE_CASES_DONE")
