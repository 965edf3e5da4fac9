"""error pattern of one fast-path case against the oracle: python tools/diag_case.py"""
import importlib, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import oracle
from test_parity_small import mode_from_spec, padding_from_spec
pkg = importlib.import_module("ndarray-conv_b200")
shape, ks, dil, mode, padding, rev = [48, 62, 350], [5, 2, 13], [1, 3, 1], ('custom', [9, 7, 2], [3, 2, 3]), ('explicit', [['replicate', 'reflect'], [('const', 1.25), 'replicate'], ['zeros', 'replicate']]), False
for cx in (False, True):
    rng = np.random.default_rng(5)
    if cx:
        x = ((rng.random(shape) - 0.5) + 1j * (rng.random(shape) - 0.5)).astype(np.complex64); k = ((rng.random(ks) - 0.5) + 1j * (rng.random(ks) - 0.5)).astype(np.complex64)
    else:
        x = (rng.random(shape, dtype=np.float32) - 0.5); k = (rng.random(ks, dtype=np.float32) - 0.5)
    ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
    kw = pkg.with_dilation(k, dil).no_reverse()
    proc = pkg.get_fft_processor(0)
    got = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
    proc.close()
    err = np.abs(got - ref); bad = err > 1e-3
    print("cx", cx, "plan", pkg.plan_query(tuple(shape), x.dtype, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding)), "max err", err.max(), "bad", int(bad.sum()), "of", bad.size)
    for ax in range(3):
        idx = np.flatnonzero(bad.any(axis=tuple(a for a in range(3) if a != ax)))
        print("   axis", ax, "bad indices:", idx[:12], "...", idx[-4:], "n", len(idx), "of", bad.shape[ax])
    w = np.argwhere(bad)[:5]
    for i in w: print("   at", tuple(i), got[tuple(i)], ref[tuple(i)])
