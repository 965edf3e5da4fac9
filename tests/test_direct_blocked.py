"""Register-blocked direct convolution (kernels_direct_tile.cuh, direct_tile_kernel<T, S2, D2>): taken for large problems only, so the
cases run in a child process with NDCONV_BLOCKED_MIN_OUT=0 -- every stride / dilation instantiation, every alignment shift of the tile,
4- and 8-byte integers and floats, zero taps, all border types -- bit for bit against the oracle (the reference's order of multiply-adds
is kept, src/conv/mod.rs:128-200)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]

# dtype, shape, kernel, dilation, mode, padding
CASES = [
    ("int32", (5, 40, 200), (3, 5, 5), 1, ("custom", [1, 2, 2], [2, 2, 2]), "replicate"),        # the BASELINE c4 geometry, small
    ("int32", (40, 200), (3, 3), 1, "same", "zeros"),
    ("int32", (3, 33, 131), (2, 3, 7), 1, "full", "reflect"),
    ("uint32", (4, 50, 150), (3, 3, 4), 2, "same", "circular"),                                  # dilation 2 on every axis
    ("int64", (4, 40, 140), (3, 3, 3), 1, ("custom", [1, 1, 1], [1, 2, 2]), "circular"),
    ("uint64", (30, 300), (5, 8), 1, "valid", "zeros"),                                          # the longest row the variant takes
    ("float32", (6, 36, 130), (3, 3, 3), 1, "same", ("const", 1.5)),
    ("float32", (64, 260), (7, 7), 1, "same", "reflect"),
    ("float32", (3, 40, 150), (2, 3, 5), [1, 1, 2], ("custom", [0, 1, 3], [1, 1, 2]), ("custom", ["zeros", "replicate", "reflect"])),   # stride 2 x dilation 2
    ("float64", (48, 200), (5, 5), 2, "same", "reflect"),
    ("float64", (4, 33, 129), (3, 3, 3), 1, "full", ("const", -0.25)),
]
# front pads 0..3 on the contiguous axis: every alignment shift of the TMA box
for pad in range(4):
    CASES.append(("int32", (20, 150), (3, 5), 1, ("explicit", [[1, 1], [pad, 2]], [1, 1]), "replicate"))
    CASES.append(("float64", (20, 150), (3, 4), 1, ("explicit", [[1, 1], [pad, 2]], [1, 2]), "reflect"))

_CHILD = r"""
import sys, importlib, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import test_direct_blocked as t
from test_parity_small import mode_from_spec, padding_from_spec
from oracle import oracle
oracle.build()
pkg = importlib.import_module("ndarray-conv_b200")
proc = pkg.get_fft_processor(0)
bad = 0
for ci, (dt, shape, ks, dil, mode, padding) in enumerate(t.CASES):
    rng = np.random.default_rng(100 + ci)
    dtype = np.dtype(dt)
    if dtype.kind in "iu":
        x = rng.integers(0 if dtype.kind == "u" else -1000, 1000, size=shape).astype(dtype); k = rng.integers(0 if dtype.kind == "u" else -9, 9, size=ks).astype(dtype)
    else:
        x = (rng.random(shape) - 0.5).astype(dtype); k = (rng.random(ks) - 0.5).astype(dtype)
        x.flat[7] = np.nan                                    # a NaN sample: only the outputs whose non-zero taps touch it may be NaN
    k.flat[1] = 0; k.flat[-2] = 0                              # zero taps: dropped by gen_offset_list, mask bits in the blocked rows
    got = pkg.conv(x, pkg.with_dilation(k, dil), mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), processor=proc)
    ref = oracle.conv(x, k, mode, padding, dil)
    names = [n for n in proc.last_kernel_names()] if hasattr(proc, "last_kernel_names") else []
    ok = got.shape == ref.shape and got.dtype == ref.dtype and np.array_equal(got, ref, equal_nan=True)
    print("CASE", ci, dt, shape, ks, "ok" if ok else "MISMATCH", flush=True)
    bad += 0 if ok else 1
print("BLOCKED_LAUNCHES", proc.launch_count)
print("RESULT", "ok" if bad == 0 else "bad")
"""


def _child(env):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", _CHILD.format(root=str(ROOT))], capture_output=True, text=True, env=e, timeout=1500)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "RESULT ok" in r.stdout, r.stdout[-3000:]
    return r


@pytest.mark.gpu
def test_blocked_direct_conv_bit_exact(pkg, cuda_lib):
    r = _child({"NDCONV_BLOCKED_MIN_OUT": "0", "NDCONV_DEBUG_BLOCKED": "1", "NDCONV_DISABLE_PERSIST": "1"})
    # the variant really ran (the library reports each blocked launch on stderr under NDCONV_DEBUG_BLOCKED)
    assert r.stderr.count("[ndconv] blocked direct conv") >= len(CASES) - 2, r.stderr[-2000:]


@pytest.mark.gpu
def test_persistent_direct_conv_bit_exact(pkg, cuda_lib):
    """the persistent double-buffered variant (large problems only by default): forced onto the small cases with THREE CTAs, so that every
    CTA walks several tiles through both window buffers, interior and edge tiles (halo patch) alike"""
    r = _child({"NDCONV_BLOCKED_MIN_OUT": "0", "NDCONV_DEBUG_BLOCKED": "1", "NDCONV_PERSIST_MIN_TILES": "0", "NDCONV_PERSIST_MAX_GRID": "3"})
    assert r.stderr.count("[ndconv] persistent direct conv") >= 3, r.stderr[-2000:]      # the stride-1 cases whose two windows fit one CTA


def test_blocked_cases_are_well_formed():
    for dt, shape, ks, dil, mode, padding in CASES:
        assert len(shape) == len(ks) and ks[-1] <= 8
