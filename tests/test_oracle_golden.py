"""Pin the CPU oracle against the reference's own known-answer vectors.

reference_kats.json : literal vectors transcribed from the reference's unit tests.
torch_golden.json   : expectations of its libtorch-derived tests, regenerated with torch CPU
                      by tests/golden/make_torch_golden.py.
"""
import json
from pathlib import Path

import numpy as np
import pytest

G = Path(__file__).parent / "golden"
KATS = json.loads((G / "reference_kats.json").read_text())
TORCH = json.loads((G / "torch_golden.json").read_text())


def _spec(p):
    """JSON lists -> the tuple specs oracle.py takes."""
    if isinstance(p, str):
        return p
    if p[0] == "const":
        return ("const", p[1])
    if p[0] == "custom":
        return ("custom", [_spec(b) for b in p[1]])
    if p[0] == "explicit":
        return ("explicit", [[_spec(b[0]), _spec(b[1])] for b in p[1]])
    raise ValueError(p)


def _mode(m):
    return m if isinstance(m, str) else (m[0], m[1], m[2])


@pytest.mark.parametrize("case", KATS["padding"], ids=lambda c: c["src"])
def test_padding_kats(oracle, case):
    x = np.array(case["x"], dtype=case["dtype"])
    got = oracle.pad(x, _spec(case["padding"]), case["pads"])
    np.testing.assert_array_equal(got, np.array(case["expect"], dtype=case["dtype"]))


@pytest.mark.parametrize("case", KATS["fft_staging"], ids=lambda c: c["src"])
def test_fft_staging_kats(oracle, case):
    if case["kind"] == "data":
        x = np.array(case["x"], dtype=case["dtype"])
        got = oracle.pad(x, _spec(case["padding"]), case["pads"], buffer_shape=case["fft_size"])
        np.testing.assert_array_equal(got, np.array(case["expect"], dtype=case["dtype"]))
    else:
        # conv_fft::padding::kernel == delta-convolution staging: check through the scipy restatement's slice rule
        k = np.array(case["kernel"], dtype=np.float64)
        d = case["dilation"]
        kd = [k.shape[i] * d[i] - d[i] + 1 for i in range(k.ndim)]
        buf = np.zeros(case["fft_size"])
        buf[tuple(slice(0, kd[i], d[i]) for i in range(k.ndim))] = k
        np.testing.assert_array_equal(buf, np.array(case["expect"], dtype=np.float64))


@pytest.mark.parametrize("case", KATS["offset_lists"], ids=lambda c: c["src"])
def test_offset_list_kats(oracle, case):
    k = np.array(case["kernel"], dtype=np.float64)
    got = oracle.gen_offset_list(k, case["dilation"], case["reverse"], case["pds_strides"])
    if "expect_len" in case:
        assert len(got) == case["expect_len"]
    else:
        assert [(o, float(w)) for o, w in got] == [(int(o), float(w)) for o, w in case["expect"]]


@pytest.mark.parametrize("case", KATS["conv"], ids=lambda c: c["src"])
def test_conv_literal_kats(oracle, case):
    x = np.array(case["x"], dtype=case["dtype"])
    k = np.array(case["kernel"], dtype=case["dtype"])
    np.testing.assert_array_equal(oracle.conv(x, k, case["mode"], case["padding"]), np.array(case["expect"], dtype=case["dtype"]))


def test_good_size_cc(oracle):
    def py(n):  # independent transliteration of good_size.rs:6-31
        b = 1
        while b < n:
            b *= 2
        while True:
            f = b // 4 * 3
            if f < n:
                break
            if f == n:
                return n
            b = f
        while True:
            f = b // 6 * 5
            if f < n:
                break
            if f == n:
                return n
            b = f
        return b
    for n, e in KATS["good_size_cc"]["pairs"]:
        assert oracle.good_size_cc(n) == e, n
    for n in list(range(1, 3000)) + [32892, 65537, 100000]:
        assert oracle.good_size_cc(n) == py(n)


@pytest.mark.parametrize("case", TORCH["int_cases"], ids=lambda c: c["src"])
def test_conv_vs_torch_golden(oracle, case):
    for dt in (np.int32, np.int64):
        x = np.array(case["x"], dtype=dt)
        k = np.array(case["kernel"], dtype=dt)
        got = oracle.conv(x, k, _mode(case["mode"]), _spec(case["padding"]), case["dilation"], case["reverse"])
        assert list(got.shape) == case["expect_shape"]
        assert got.ravel().tolist() == case["expect"]


@pytest.mark.parametrize("case", TORCH["int_cases"], ids=lambda c: c["src"])
def test_conv_fft_restatement_vs_torch_golden(oracle, case):
    # the reference's own gate: round(conv_fft) == conv, tol 1e-5 (f32) / 1e-9 (f64)  (conv_fft/tests.rs:15-16)
    for dt, tol in ((np.float32, 1e-5), (np.float64, 1e-9)):
        x = np.array(case["x"], dtype=dt)
        k = np.array(case["kernel"], dtype=dt)
        got = oracle.conv_fft(x, k, _mode(case["mode"]), _spec(case["padding"]), case["dilation"], case["reverse"])
        assert list(got.shape) == case["expect_shape"]
        assert np.max(np.abs(np.rint(got.ravel()) - np.array(case["expect"]))) < tol
        sp = oracle.conv_fft_scipy(x, k, _mode(case["mode"]), _spec(case["padding"]), case["dilation"], case["reverse"])
        assert np.max(np.abs(np.rint(sp.ravel()) - np.array(case["expect"]))) < tol


@pytest.mark.parametrize("case", TORCH["float_cases"], ids=lambda c: c["src"])
def test_float_circular_kat(oracle, case):
    x = np.array(case["x"], np.float32)
    k = np.array(case["kernel"], np.float32)
    e = np.array(case["expect"])
    assert np.max(np.abs(oracle.conv(x, k, case["mode"], case["padding"]) - e)) < case["tol"]
    assert np.max(np.abs(oracle.conv_fft(x, k, case["mode"], case["padding"]) - e)) < case["tol"]


def test_sequential_vs_closed_form_padding(oracle):
    """the literal sequential restatement and the closed form agree on the well-defined domain."""
    rng = np.random.default_rng(7)
    borders = ["zeros", ("const", 5), "reflect", "replicate", "circular"]
    n_checked = 0
    for _ in range(400):
        nd = int(rng.integers(1, 4))
        shape = [int(rng.integers(1, 6)) for _ in range(nd)]
        x = rng.integers(-9, 10, size=shape).astype(np.int32)
        pads, spec = [], []
        for i in range(nd):
            sides = []
            pp = []
            for s in range(2):
                b = borders[int(rng.integers(0, 5))]
                hi = {"reflect": shape[i] - 1, "circular": shape[i] if s == 0 else 7}.get(b if isinstance(b, str) else "", 7)
                pp.append(int(rng.integers(0, hi + 1)))
                sides.append(b)
            pads.append(pp)
            spec.append(sides)
        a = oracle.pad(x, ("explicit", spec), pads)
        b = oracle.pad(x, ("explicit", spec), pads, closed_form=True)
        np.testing.assert_array_equal(a, b)
        n_checked += 1
    assert n_checked == 400


def test_padding_vs_torch_pad(oracle):
    """src/padding/mod.rs:763-861 (aligned_with_libtorch): 5 modes x 1/2/3-D against F.pad."""
    torch = pytest.importorskip("torch")
    import torch.nn.functional as F
    rng = np.random.default_rng(3)
    for nd in (1, 2, 3):
        shape = [4, 5, 6][:nd]
        x = rng.integers(0, 100, size=shape).astype(np.float64)
        pads = [[1, 2], [2, 1], [3, 3]][:nd]
        flat = []
        for p in reversed(pads):
            flat += p
        t = torch.from_numpy(x)[None, None]
        for name, tm in (("zeros", "constant"), ("reflect", "reflect"), ("replicate", "replicate"), ("circular", "circular")):
            e = (F.pad(t, flat, mode=tm) if tm != "constant" else F.pad(t, flat, value=0.0))[0, 0].numpy()
            np.testing.assert_array_equal(oracle.pad(x, name, pads), e)
        e = F.pad(t, flat, value=7.0)[0, 0].numpy()
        np.testing.assert_array_equal(oracle.pad(x, ("const", 7), pads), e)


def test_errors(oracle):
    from oracle.oracle import OracleError, DATA_SHAPE, KERNEL_SHAPE, MISMATCH_SHAPE
    x = np.zeros((0, 3), np.int32)
    k = np.ones((1, 1), np.int32)
    with pytest.raises(OracleError) as e:
        oracle.conv(x, k)
    assert e.value.status == DATA_SHAPE
    with pytest.raises(OracleError) as e:
        oracle.conv(np.ones((2, 2), np.int32), np.ones((0, 1), np.int32))
    assert e.value.status == KERNEL_SHAPE
    with pytest.raises(OracleError) as e:  # conv_fft quirk: DataShape for an empty kernel (conv_fft/mod.rs:211-213)
        oracle.conv_fft(np.ones((2, 2), np.float32), np.ones((0, 1), np.float32))
    assert e.value.status == DATA_SHAPE
    with pytest.raises(OracleError) as e:
        oracle.conv(np.ones(3, np.int32), np.ones(5, np.int32), "valid")
    assert e.value.status == MISMATCH_SHAPE


def test_direct_vs_fft_random(oracle):
    rng = np.random.default_rng(11)
    for trial in range(40):
        nd = int(rng.integers(1, 4))
        shape = [int(rng.integers(3, 9)) for _ in range(nd)]
        ks = [int(rng.integers(1, 4)) for _ in range(nd)]
        d = int(rng.integers(1, 3))
        x = rng.random(shape)
        k = rng.random(ks)
        mode = ["full", "same", "valid"][trial % 3]
        if mode == "valid" and any((kk - 1) * d + 1 > s for kk, s in zip(ks, shape)):
            continue
        pad = ["zeros", "replicate", "circular", ("const", 0.5)][trial % 4]
        rev = bool(trial % 2)
        a = oracle.conv(x, k, mode, pad, d, rev)
        b = oracle.conv_fft(x, k, mode, pad, d, rev)
        c = oracle.conv_fft_scipy(x, k, mode, pad, d, rev)
        assert np.max(np.abs(a - b)) < 1e-10
        assert np.max(np.abs(a - c)) < 1e-10
    # complex
    x = rng.random((5, 6)) + 1j * rng.random((5, 6))
    k = rng.random((2, 3)) + 1j * rng.random((2, 3))
    a = oracle.conv(x, k, "same", "reflect")
    b = oracle.conv_fft(x, k, "same", "reflect")
    c = oracle.conv_fft_scipy(x, k, "same", "reflect")
    assert np.max(np.abs(a - b)) < 1e-10 and np.max(np.abs(a - c)) < 1e-10
