"""premise check: do several host threads (handles on ONE device) speed up host-resident conv_fft on PAGEABLE arrays?"""
import importlib, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("ndarray-conv_b200")
lib = pkg.get_library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(0)
k = rng.random((63, 63), dtype=np.float32)
x = rng.random((n, n), dtype=np.float32)
out = np.zeros((n + 62, n + 62), np.float32)
for nh in (1, 2, 3, 4, 6):
    procs = [pkg.get_fft_processor(0, lib) for _ in range(nh)]
    pkg.conv_fft_sharded(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs, out=out)
    t0 = time.perf_counter()
    for _ in range(2):
        pkg.conv_fft_sharded(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs, out=out)
    t = (time.perf_counter() - t0) / 2
    print(f"pageable, {nh} handle(s) on one GPU: {t * 1e3:.1f} ms, {(x.nbytes + out.nbytes) / t / 1e9:.1f} GB/s aggregate", flush=True)
    for p in procs:
        p.close()
