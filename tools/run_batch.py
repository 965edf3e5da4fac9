"""Throughput of ndconv_conv_fft_batch on small independent problems (host-resident c2- and c3-shaped inputs) against the number of
processor handles on one GPU.  usage: python tools/run_batch.py"""
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("ndarray-conv_b200")
lib = pkg.get_library()
rng = np.random.default_rng(0)
B = pkg.BorderType
CASES = [("c2", (200, 5000), (11, 31), 2, pkg.ConvMode.Same, pkg.PaddingMode.Custom([B.Reflect, B.Circular])),
         ("c3", (10, 100, 200), (5, 11, 31), 1, pkg.ConvMode.Same, pkg.PaddingMode.Zeros)]
for name, xs, ks, dil, mode, pm in CASES:
    n = 64
    xb = [rng.random(xs, dtype=np.float32) for _ in range(n)]
    k = pkg.with_dilation(rng.random(ks, dtype=np.float32), dil)
    ref = None
    for nh in (1, 2, 4, 8):
        procs = [pkg.get_fft_processor(0, lib) for _ in range(nh)]
        outs = pkg.conv_fft_batch(xb, k, mode, pm, procs)      # warm
        t0 = time.perf_counter()
        for _ in range(3):
            outs = pkg.conv_fft_batch(xb, k, mode, pm, procs)
        t = (time.perf_counter() - t0) / 3
        if ref is None:
            ref = outs
        same = all(np.array_equal(a, b) for a, b in zip(ref, outs))
        print(json.dumps({"shape": name, "problems": n, "handles": nh, "ms_per_batch": round(t * 1e3, 2), "us_per_problem": round(t * 1e6 / n, 1), "identical_to_1_handle": same}), flush=True)
        for p in procs:
            p.close()

# ---- device-resident: problems are only enqueued (one stream per handle); the problem array is built once, as a native caller would ----
import ctypes  # noqa: E402
import torch  # noqa: E402
dev = torch.device("cuda", 0)
for name, xs, ks, dil, mode, pm in CASES:
    n = 64
    xd = [torch.rand(xs, device=dev) for _ in range(n)]
    k = pkg.with_dilation(rng.random(ks, dtype=np.float32), dil)
    strides = [int(np.prod(xs[i + 1:])) for i in range(len(xs))]
    prs, keeps = (pkg._Problem * n)(), []
    for i in range(n):
        pr, keep = pkg.make_problem(xs, strides, xd[i].data_ptr(), np.float32, k, mode, pm, pkg.MEM_DEVICE, lib)
        prs[i] = pr
        keeps.append(keep)
    oshape = pkg.out_shape(prs[0], pkg.PATH_FFT, lib)
    yd = [torch.empty(tuple(oshape), device=dev) for _ in range(n)]
    optr = (ctypes.c_void_p * n)(*[y.data_ptr() for y in yd])
    ref = None
    for nh in (1, 2, 4, 8):
        procs = [pkg.get_fft_processor(0, lib) for _ in range(nh)]
        handles = (ctypes.c_void_p * nh)(*[p.handle for p in procs])
        def run():
            lib.check(lib.c.ndconv_conv_fft_batch(handles, nh, prs, optr, n))
            for p in procs:
                p.synchronize()
        run()
        t0 = time.perf_counter()
        for _ in range(5):
            run()
        t = (time.perf_counter() - t0) / 5
        got = torch.stack(yd).cpu().numpy()
        if ref is None:
            ref = got
        print(json.dumps({"shape": name + " device-resident", "problems": n, "handles": nh, "us_per_problem": round(t * 1e6 / n, 2),
                          "identical_to_1_handle": bool(np.array_equal(ref, got))}), flush=True)
        for p in procs:
            p.close()
