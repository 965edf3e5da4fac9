"""Time ndconv_conv_fft_sharded (one process, one host thread per GPU) on the c5 workload with pinned host buffers, for
1..N processor handles on devices 0..N-1.  usage: python tools/run_sharded.py [rows=32768] [steps=3]
Prints one JSON object per handle count: e2e ms / Gsamples/s, and the max |difference| against the single-handle result."""
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("ndarray-conv_b200")


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    lib = pkg.get_library()
    ndev = lib.c.ndconv_device_count()
    n1, kk = 32768, 63
    x = torch.empty((rows, n1), dtype=torch.float32).pin_memory()
    g = torch.Generator().manual_seed(1004)
    x.uniform_(0, 1, generator=g)
    out = torch.empty((rows + kk - 1, n1 + kk - 1), dtype=torch.float32).pin_memory()
    k = np.random.default_rng(2004).random((kk, kk), dtype=np.float32)
    xn, on = x.numpy(), out.numpy()
    ref = None
    counts = sorted({1, 2, ndev} & set(range(1, ndev + 1))) if ndev > 1 else [1, 2]
    for n in counts:
        procs = [pkg.get_fft_processor(d % ndev, lib) for d in range(n)]
        pkg.conv_fft_sharded(xn, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs, out=on)     # warm: plans, kernel spectra, staging buffers
        t0 = time.perf_counter()
        for _ in range(steps):
            pkg.conv_fft_sharded(xn, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs, out=on)
        t = (time.perf_counter() - t0) / steps
        if ref is None:
            ref = on[::97].copy()
            diff = 0.0
        else:
            diff = float(np.max(np.abs(on[::97] - ref)))
        print(json.dumps({"handles": n, "devices": min(n, ndev), "e2e_ms": round(t * 1e3, 2), "Gsamples_per_s": round(on.size / t / 1e9, 2),
                          "max_abs_diff_vs_1": diff, "launches": [p.launch_count for p in procs]}), flush=True)
        for p in procs:
            p.close()


if __name__ == "__main__":
    main()
