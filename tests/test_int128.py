"""i128 / u128 elements of the direct `conv` (T: NumAssign + Copy, src/conv/mod.rs:118-121) and the zero-tap
semantics of gen_offset_list (src/dilation/mod.rs:49: a 0 weight is dropped from the tap list, so it never multiplies a NaN / Inf
sample).  numpy has no 128-bit integer: arrays are (lo, hi) records (pkg.I128 / pkg.U128), Python ints are the second opinion."""
import itertools

import numpy as np
import pytest

from test_parity_small import mode_from_spec, padding_from_spec

M = 1 << 128


def bigint_conv(x, k, pads, strides, dil, reverse, border_fn):
    """out[o] = sum_j B[o s + j d] kf[j] mod 2^128 with Python ints; border_fn(axis, i, n, pf) -> source index or a constant"""
    nd = x.ndim
    kd = [(k.shape[a] - 1) * dil[a] + 1 for a in range(nd)]
    P = [x.shape[a] + pads[a][0] + pads[a][1] for a in range(nd)]
    O = [(P[a] - kd[a]) // strides[a] + 1 for a in range(nd)]
    out = np.zeros(O, dtype=object)
    for o in itertools.product(*[range(v) for v in O]):
        acc = 0
        for j in itertools.product(*[range(v) for v in k.shape]):
            w = k[tuple(k.shape[a] - 1 - j[a] for a in range(nd))] if reverse else k[j]
            if w == 0:
                continue
            src, const = [], None
            for a in range(nd):
                r = border_fn(a, o[a] * strides[a] + j[a] * dil[a], x.shape[a], pads[a][0])
                if isinstance(r, tuple):
                    const = r[0]
                src.append(r)
            v = const if const is not None else x[tuple(src)]
            acc = (acc + v * w) % M
        out[o] = acc
    return out


def replicate(a, i, n, pf):
    return min(max(i - pf, 0), n - 1)


def zeros(a, i, n, pf):
    return i - pf if 0 <= i - pf < n else (0,)


CASES = [
    # shape, kernel shape, pads, strides, dilation, reverse, border name, border fn, signed
    ((9,), (3,), [[1, 1]], [1], [1], True, "replicate", replicate, True),
    ((5, 7), (2, 3), [[1, 0], [2, 2]], [2, 1], [1, 2], False, "zeros", zeros, False),
    ((3, 4, 5), (2, 2, 3), [[1, 1], [0, 1], [1, 1]], [1, 2, 1], [1, 1, 1], True, "replicate", replicate, True),
]


def make(rng, shape, signed, big=True):
    vals = np.empty(shape, dtype=object)
    for idx in np.ndindex(*shape):
        mag = int(rng.integers(0, 1 << 62)) * (int(rng.integers(0, 1 << 62)) if big else 1)
        vals[idx] = -mag if (signed and rng.integers(0, 2)) else mag
    return vals


@pytest.mark.parametrize("case", CASES, ids=[str(c[0]) for c in CASES])
def test_oracle_matches_python_bigints(pkg, oracle, case):
    shape, ks, pads, strides, dil, rev, bname, bfn, signed = case
    rng = np.random.default_rng(7)
    xv, kv = make(rng, shape, signed), make(rng, ks, signed)
    kv[(0,) * len(ks)] = 0                                   # a dropped tap
    x, k = pkg.int128_array(xv, signed), pkg.int128_array(kv, signed)
    got = oracle.conv(x, k, ("explicit", pads, strides), bname, dil, rev)
    want = bigint_conv(xv % M, kv % M, pads, strides, dil, rev, bfn)
    assert got.shape == want.shape
    assert np.array_equal(pkg.int128_values(got) % M, want)   # wrapping products (|x| |w| ~ 2^248) included


@pytest.mark.parametrize("signed", [True, False])
def test_int128_conv_bit_exact(ndc, oracle, signed):
    pkg, lib = ndc                                     # kernel bodies under host emulation on the CPU box, the sm_100a build on a B200
    rng = np.random.default_rng(11)
    proc = pkg.get_fft_processor(0, lib)
    cases = [((40,), (5,), 1, "same", "reflect", True), ((16, 40, 48), (3, 5, 5), 1, ("custom", [1, 2, 2], [2, 2, 2]), "replicate", True),
             ((9, 30, 50), (2, 3, 3), 2, "full", ("const", -3), False), ((3, 4, 5, 6), (2, 2, 2, 2), 1, "same", "circular", True)]
    for shape, ks, dil, mode, padding, rev in cases:
        x, k = pkg.int128_array(make(rng, shape, signed), signed), pkg.int128_array(make(rng, ks, signed), signed)
        kw = pkg.with_dilation(k, dil)
        if not rev:
            kw = kw.no_reverse()
        got = pkg.conv(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), processor=proc)
        ref = oracle.conv(x, k, mode, padding, dil, rev)
        assert got.dtype == x.dtype and got.shape == ref.shape
        assert got.tobytes() == ref.tobytes(), (shape, signed)
    proc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64])
def test_zero_weight_tap_never_touches_nan(pkg, cuda_lib, oracle, dt):
    """src/dilation/mod.rs:49: zero weights are filtered out of the offset list, so NaN / Inf samples under a zero tap do not
    poison the output (0 * NaN would).  Both the tile kernel (rank 3) and the generic kernel (rank 4) are covered."""
    rng = np.random.default_rng(3)
    proc = pkg.get_fft_processor(0, cuda_lib)
    for shape, ks in (((12, 20, 33), (3, 3, 3)), ((4, 5, 6, 7), (1, 3, 1, 3))):
        x = rng.random(shape).astype(dt)
        k = (rng.random(ks) + 0.5).astype(dt)
        k[..., 1] = 0                                    # the middle tap column of the last axis is zero ...
        xs = x
        xs[..., 1::3] = np.nan                           # ... and the samples 1, 4, 7, ... along it are NaN / Inf:
        xs[..., 1] = np.inf if dt != np.complex64 else complex(np.inf, -np.inf)
        # no padding, stride 3 on the last axis, no_reverse: output o reads samples 3o, 3o+1, 3o+2 with weights k[0], 0, k[2]
        strides = [1] * (len(shape) - 1) + [3]
        got = pkg.conv(xs, pkg.with_dilation(k, 1).no_reverse(), pkg.ConvMode.Custom([0] * len(shape), strides), pkg.PaddingMode.Zeros, processor=proc)
        ref = oracle.conv(xs, k, ("custom", [0] * len(shape), strides), "zeros", 1, False)
        assert np.isfinite(ref).all() and got.tobytes() == ref.tobytes()
    proc.close()
