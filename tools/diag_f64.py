"""where does the f64 fast path differ from the oracle?  python tools/diag_f64.py"""
import importlib, sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle
pkg = importlib.import_module("ndarray-conv_b200")
rng = np.random.default_rng(1)
def run(shape, k, mode="same", pad="zeros", tag=""):
    x = rng.random(shape) - 0.25
    proc = pkg.get_fft_processor(0)
    got = pkg.conv_fft_with_processor(x, k, getattr(pkg.ConvMode, mode.capitalize()), getattr(pkg.PaddingMode, pad.capitalize()), proc)
    proc.close()
    ref = oracle.conv_f64_truth(x, k, mode, pad, 1, True)
    err = np.abs(got - ref)
    bad = err > 1e-9
    info = pkg.plan_query(shape, np.float64, k, getattr(pkg.ConvMode, mode.capitalize()), getattr(pkg.PaddingMode, pad.capitalize()))
    print(tag, shape, k.shape, "tiles", info["tile_len"], info["n_tiles"], "max err %.3e" % err.max(), "bad cells %d of %d" % (bad.sum(), bad.size))
    if bad.any():
        rows = np.flatnonzero(bad.any(axis=tuple(range(1, bad.ndim)))); cols = np.flatnonzero(bad.any(axis=tuple(range(0, bad.ndim - 1))))
        print("   bad rows: n=%d first %s last %s | bad last-axis cols: n=%d first %s last %s" % (len(rows), rows[:6], rows[-3:], len(cols), cols[:8], cols[-3:]))
        w = np.argwhere(bad)[:4]
        for idx in w: print("   at", tuple(idx), "got", got[tuple(idx)], "ref", ref[tuple(idx)], "ratio", got[tuple(idx)] / ref[tuple(idx)] if ref[tuple(idx)] else None)
delta = np.ones((1, 1))
run((140, 150), delta, tag="delta T=8?")
run((140, 100), delta, tag="delta 128 row")
run((140, 400), delta, tag="delta 512 row")
run((40, 400), delta, tag="delta short cols")
run((300, 400), delta, tag="delta 2 col tiles?")
run((140, 400), np.array([[0.0, 0.0, 1.0]]), tag="row shift")
run((140, 400), np.array([[0.0], [0.0], [1.0]]), tag="col shift")
run((140, 400), rng.random((1, 3)), tag="row k3")
run((140, 400), rng.random((3, 1)), tag="col k3")
run((700, 1500), rng.random((5, 9)) - 0.5, tag="case0")
