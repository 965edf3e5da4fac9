"""col_pass_tma_kres (kernels_fft_fast.cuh): the 1024-row column pass with the kernel spectrum resident in Tensor Memory.  It runs when
axis 0 has at least 10 tiles of 1024 rows; the cases cover the geometry of its work order -- one bundle only, a long last bundle, fewer
chunks than CTAs, rank 3 (an inner extent of several rows) -- against the f64 oracle, and bit for bit against col_pass_tma (the same
butterflies fed from L2; NDCONV_DISABLE_COL_KRES=1 in a subprocess)."""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from test_parity_small import fft_tol, mode_from_spec, padding_from_spec

ROOT = Path(__file__).resolve().parents[1]
CASES = [
    # shape, kernel, dilation, mode, padding, reverse
    ((4000, 6500), (9, 5), 1, "full", "reflect", True),                                       # 4 x 13 tiles, 33 blocks
    ((12000, 1000), (5, 5), 1, "same", ("custom", ["replicate", "circular"]), False),         # 12 tiles: one bundle
    ((20000, 300), (63, 3), 1, "full", "zeros", True),                                        # 21 tiles, 63-row kernel (skip = 62)
    ((10200, 520), (3, 9), 1, "same", ("const", 0.5), True),                                  # 30 tiles x 17 blocks: fewer chunks than CTAs
    ((8192, 9868), (63, 63), 1, "full", "reflect", True),                                     # the reduced BASELINE workload (row pitch 1032)
    ((1000, 6, 11000), (3, 3, 3), 1, "full", "replicate", True),                               # rank 3: 11 tiles along the last axis, inner = F1 x pitch (1040 blocks)
]


def planned_for_kres(pkg, shape, ks, dil, mode, padding):
    info = pkg.plan_query(shape, np.float32, pkg.with_dilation(np.ones(ks, np.float32), dil), mode_from_spec(pkg, mode), padding_from_spec(pkg, padding))
    ntiles = int(np.prod(info["n_tiles"]))
    return info["path"] == "fast" and info["tile_len"][0] == 1024 and ntiles >= 10, info


@pytest.mark.parametrize("case", CASES, ids=[str(c[:2]) for c in CASES])
def test_cases_take_the_resident_path(pkg, case):
    """host logic only: the planner gives these shapes 1024-row axis-0 tiles and at least 10 tiles (otherwise the GPU cases below test nothing)"""
    shape, ks, dil, mode, padding, rev = case
    ok, info = planned_for_kres(pkg, shape, ks, dil, mode, padding)
    assert ok, info


def _inputs(shape, ks):
    rng = np.random.default_rng(7)
    return rng.random(shape, dtype=np.float32) - 0.25, rng.random(ks, dtype=np.float32) - 0.5


def _run(pkg, case):
    shape, ks, dil, mode, padding, rev = case
    x, k = _inputs(shape, ks)
    kw = pkg.with_dilation(k, dil)
    if not rev:
        kw = kw.no_reverse()
    proc = pkg.get_fft_processor(0)
    try:
        got = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
        again = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
        names = [n for n, _ in proc.kernel_names()] if hasattr(proc, "kernel_names") else []
    finally:
        proc.close()
    return x, k, got, again, names


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[str(c[:2]) for c in CASES])
def test_resident_vs_oracle(pkg, cuda_lib, oracle, case):
    shape, ks, dil, mode, padding, rev = case
    x, k, got, again, _ = _run(pkg, case)
    assert np.array_equal(got, again)                     # second call: cached kernel spectrum and kres layout
    # the f64 truth on output bands (whole arrays of 80 M samples x 4000 taps would take minutes): the first rows, a band across
    # an axis-0 tile seam, the last rows
    ref = None
    if np.prod(shape) * np.prod(ks) <= 3e9:
        ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
        assert got.shape == ref.shape
        tol = fft_tol(np.float32, 1024 * 2048, ref, float(np.max(np.abs(x)) * np.sum(np.abs(k))))
        err = float(np.max(np.abs(got - ref)))
        assert err <= tol, (err, tol)
    else:
        assert np.isfinite(got).all()
        # a full-height column band is a complete problem in its own right along axis 0 (all axis-0 tiles, every bundle) and cheap for
        # the f64 truth: Full mode with a Reflect border, so the band's interior columns do not depend on what lies beside it
        c0, w = 5000, 64
        kh = ks[1]
        xb = x[:, c0 - kh: c0 + w + kh]
        refb = oracle.conv_f64_truth(xb, k, mode, padding, dil, rev)
        # Full mode: output column o reads input columns [o - (kh - 1), o]; band column q (output) = global column c0 - kh + q
        sub = refb[:, 2 * kh: 2 * kh + w]
        mine = got[:, c0 + kh: c0 + kh + w]
        tol = fft_tol(np.float32, 1024 * 2048, sub, float(np.max(np.abs(x)) * np.sum(np.abs(k))))
        err = float(np.max(np.abs(mine - sub)))
        assert err <= tol, (err, tol)


_CHILD = r"""
import sys, hashlib, importlib, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import test_col_kres as t
pkg = importlib.import_module("ndarray-conv_b200")
for case in t.CASES:
    x, k, got, again, _ = t._run(pkg, case)
    print(hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest())
"""


@pytest.mark.gpu
def test_resident_equals_l2_fed_kernel_bit_for_bit(pkg, cuda_lib):
    """same butterflies, same operation order: only where the factors come from differs"""
    outs = {}
    for tag, env in (("kres", {}), ("l2", {"NDCONV_DISABLE_COL_KRES": "1"})):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", _CHILD.format(root=str(ROOT))], capture_output=True, text=True, env=e, timeout=1500)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = r.stdout.split()
    assert len(outs["kres"]) == len(CASES)
    assert outs["kres"] == outs["l2"]
