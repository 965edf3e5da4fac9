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
