"""Repeat one fast-path case with fresh processors while the device memory they will allocate is poisoned with NaN first (a freed
torch block of NaN is handed back to the driver, so the next cudaMalloc'ed workspace starts as NaN): catches reads of workspace
cells no earlier pass has written, and races that only show under some timing.  usage: python tools/stress_case.py [reps]"""
import importlib, sys
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import oracle
pkg = importlib.import_module("ndarray-conv_b200")
from test_parity_small import fft_tol, mode_from_spec, padding_from_spec

CASES = [((1000, 20, 300), (3, 3, 3), 1, "full", ("custom", ["replicate", "circular", "zeros"]), False),
         ((3, 1000, 300), (2, 3, 5), 1, "same", "reflect", True),
         ((1300, 2600), (5, 9), 1, "full", "reflect", True),
         # round 2b: the Tensor-Memory-resident column pass (>= 10 tiles of 1024 rows): few chunks per CTA / one bundle / the c5 tile shape
         ((10200, 520), (3, 9), 1, "same", ("const", 0.5), True),
         ((12000, 1000), (5, 5), 1, "same", ("custom", ["replicate", "circular"]), False),
         ((4000, 6500), (9, 5), 1, "full", "reflect", True)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
bad = 0
for shape, ks, dil, mode, padding, rev in CASES:
    rng = np.random.default_rng(42)
    x = rng.random(shape, dtype=np.float32) - 0.25
    k = rng.random(ks, dtype=np.float32) - 0.5
    kw = pkg.with_dilation(k, dil)
    if not rev:
        kw = kw.no_reverse()
    ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
    tol = fft_tol(np.float32, 1024 * 2048, ref, float(np.max(np.abs(x)) * np.sum(np.abs(k))))
    for r in range(reps):
        junk = torch.full((1 << 28,), float("nan") if r % 2 == 0 else 1e30, device="cuda")   # 1 GiB of poison
        del junk
        torch.cuda.synchronize(); torch.cuda.empty_cache()
        proc = pkg.get_fft_processor(0)
        got = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
        got2 = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
        err = float(np.nanmax(np.abs(got - ref))) if np.isfinite(got).any() else float("inf")
        ok = np.isfinite(got).all() and err <= tol and np.array_equal(got, got2)
        if not ok:
            bad += 1
            w = np.argwhere(~np.isfinite(got) | (np.abs(got - ref) > tol))
            print("FAIL", shape, "rep", r, "err", err, "tol", tol, "same twice", np.array_equal(got, got2), "bad cells", len(w), "first", w[:3].tolist(), "last", w[-3:].tolist())
        proc.close()
print("STRESS_DONE failures:", bad)
