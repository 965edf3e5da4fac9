"""experiment: host-side cost of one PreparedConv call (enqueue rate) against the device time per call, c1 (1-D 5000, k=31)."""
import importlib, sys, time
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
pkg = importlib.import_module("ndarray-conv_b200")
lib = pkg.get_library()
dev = torch.device("cuda", 0)
proc = pkg.get_fft_processor(0, lib)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
x = torch.rand(5000, device=dev)
k = np.random.default_rng(0).random(31, dtype=np.float32)
prep = pkg.PreparedConv("ndconv_conv_fft", proc, (5000,), (1,), np.float32, pkg.with_dilation(k, 1), pkg.ConvMode.Same, pkg.PaddingMode.Zeros)
y = torch.empty(prep.out_shape, device=dev)
xp, yp = x.data_ptr(), y.data_ptr()
for _ in range(20):
    prep(xp, yp)
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
for _ in range(n):
    prep(xp, yp)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e6 * (t1 - t0) / n:.2f} us/call (host), drained after {1e6 * (t2 - t0) / n:.2f} us/call (host + device)")
# the same launches replayed from a CUDA graph: device time per call without the host in the loop
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=st):
    for _ in range(100):
        prep(xp, yp)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
print(f"CUDA graph of 100 calls: {e0.elapsed_time(e1) * 10:.2f} us/call (device only)")
