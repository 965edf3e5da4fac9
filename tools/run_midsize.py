"""Per-call time of mid-size 2-D / 3-D f32 conv_fft problems (device-resident, warm processor) -- used to place the planner's
"small problem" threshold (NDCONV_TILE_SMALL_K / NDCONV_TILE_SLACK).  usage: [env ...] python tools/run_midsize.py"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

pkg = importlib.import_module("ndarray-conv_b200")
lib = pkg.get_library()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
proc = pkg.get_fft_processor(0, lib)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
proc.set_stream(stream.cuda_stream)
CASES = [((700, 700), (15, 15)), ((1000, 1000), (15, 15)), ((1400, 1400), (31, 31)), ((2000, 2000), (31, 31)), ((3000, 3000), (31, 31)),
         ((4096, 4096), (63, 63)), ((8192, 8192), (63, 63)), ((64, 200, 200), (5, 7, 7)), ((128, 300, 400), (5, 7, 7))]
res = []
for xs, ks in CASES:
    rng = np.random.default_rng(1)
    xd = torch.from_numpy(rng.random(xs, dtype=np.float32)).to(dev)
    kh = rng.random(ks, dtype=np.float32)
    strides = [int(np.prod(xs[i + 1:])) for i in range(len(xs))]
    prep = pkg.PreparedConv("ndconv_conv_fft", proc, xs, strides, np.float32, pkg.with_dilation(kh, 1), pkg.ConvMode.Same, pkg.PaddingMode.Reflect)
    yd = torch.empty(int(np.prod(prep.out_shape)), dtype=torch.float32, device=dev)
    xp, yp = xd.data_ptr(), yd.data_ptr()
    for _ in range(3):
        prep(xp, yp)
    torch.cuda.synchronize(dev)
    reps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        prep(xp, yp)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    res.append(("x".join(map(str, xs)), round(e0.elapsed_time(e1) * 1e3 / reps, 1)))
print(res)
proc.close()
