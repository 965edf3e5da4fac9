"""Parity of the product against the CPU oracle on small seeded cases and the reference's KATs.

Each test runs twice: `emul` (kernel bodies compiled as host C++, CPU box -- checks index logic without a GPU;
test infrastructure, never the product) and `cuda` (-m gpu: the real sm_100a library through the C ABI)."""
import json
from pathlib import Path

import numpy as np
import pytest

G = Path(__file__).parent / "golden"
KATS = json.loads((G / "reference_kats.json").read_text())
TORCH = json.loads((G / "torch_golden.json").read_text())


def mode_from_spec(pkg, m):
    if isinstance(m, str):
        return {"full": pkg.ConvMode.Full, "same": pkg.ConvMode.Same, "valid": pkg.ConvMode.Valid}[m]
    return pkg.ConvMode.Custom(m[1], m[2]) if m[0] == "custom" else pkg.ConvMode.Explicit(m[1], m[2])


def border_from_spec(pkg, b):
    if isinstance(b, str):
        return {"zeros": pkg.BorderType.Zeros, "reflect": pkg.BorderType.Reflect, "replicate": pkg.BorderType.Replicate,
                "circular": pkg.BorderType.Circular}[b]
    return pkg.BorderType.Const(b[1])


def padding_from_spec(pkg, p):
    if isinstance(p, str):
        return {"zeros": pkg.PaddingMode.Zeros, "reflect": pkg.PaddingMode.Reflect, "replicate": pkg.PaddingMode.Replicate,
                "circular": pkg.PaddingMode.Circular}[p]
    if p[0] == "const":
        return pkg.PaddingMode.Const(p[1])
    if p[0] == "custom":
        return pkg.PaddingMode.Custom([border_from_spec(pkg, b) for b in p[1]])
    return pkg.PaddingMode.Explicit([[border_from_spec(pkg, b[0]), border_from_spec(pkg, b[1])] for b in p[1]])


def fft_tol(dtype, fft_points, ref, bound=None):
    """|err| <= c * eps * log2(N) * max|out|, c = 4.  `bound` (= max|x| * sum|k| >= max|out|) is only
    used when the exact output is identically ~0 (every tap lands in a zero border), where max|out| gives no scale."""
    eps = np.finfo(np.float32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64).eps
    scale = float(np.max(np.abs(ref)))
    if bound is not None and scale < 1e-9 * bound:
        scale = bound
    return 4.0 * eps * max(np.log2(max(fft_points, 2)), 1.0) * max(scale, 1e-30)


@pytest.mark.parametrize("case", TORCH["int_cases"], ids=lambda c: c["src"])
def test_conv_direct_reference_cases(ndc, case):
    pkg, lib = ndc
    for dt in (np.int32, np.int64, np.float32):
        x = np.array(case["x"], dtype=dt)
        k = pkg.with_dilation(np.array(case["kernel"], dtype=dt), case["dilation"])
        if not case["reverse"]:
            k = k.no_reverse()
        got = pkg.conv(x, k, mode_from_spec(pkg, case["mode"]), padding_from_spec(pkg, case["padding"]), lib=lib)
        assert list(got.shape) == case["expect_shape"]
        assert got.ravel().tolist() == case["expect"]


@pytest.mark.parametrize("case", TORCH["int_cases"], ids=lambda c: c["src"])
def test_conv_fft_reference_cases(ndc, case):
    # the reference's own gate (src/conv_fft/tests.rs:15-16,37-39): round(conv_fft) == conv, 1e-5 f32 / 1e-9 f64
    pkg, lib = ndc
    for dt, tol in ((np.float32, 1e-5), (np.float64, 1e-9)):
        x = np.array(case["x"], dtype=dt)
        k = pkg.with_dilation(np.array(case["kernel"], dtype=dt), case["dilation"])
        if not case["reverse"]:
            k = k.no_reverse()
        got = pkg.conv_fft(x, k, mode_from_spec(pkg, case["mode"]), padding_from_spec(pkg, case["padding"]), lib=lib)
        assert list(got.shape) == case["expect_shape"]
        e = np.array(case["expect"], dtype=np.float64)
        assert np.max(np.abs(np.rint(got.ravel()) - e)) < tol
        assert np.max(np.abs(got.ravel() - e)) < 64 * np.finfo(dt).eps * max(np.max(np.abs(e)), 1)


@pytest.mark.parametrize("case", KATS["conv"], ids=lambda c: c["src"])
def test_conv_literal_kats(ndc, case):
    pkg, lib = ndc
    x = np.array(case["x"], dtype=case["dtype"])
    k = np.array(case["kernel"], dtype=case["dtype"])
    got = pkg.conv(x, k, mode_from_spec(pkg, case["mode"]), padding_from_spec(pkg, case["padding"]), lib=lib)
    np.testing.assert_array_equal(got, np.array(case["expect"], dtype=case["dtype"]))


@pytest.mark.parametrize("case", KATS["padding"], ids=lambda c: c["src"])
def test_padding_kats_through_identity_conv(ndc, case):
    """The padded buffer is never materialised; convolving with the 1-tap identity kernel under
    ConvMode::Explicit{pads} exposes exactly the padded array the reference's `padding` would build."""
    pkg, lib = ndc
    x = np.array(case["x"], dtype=case["dtype"])
    k = np.ones([1] * x.ndim, dtype=case["dtype"])
    mode = pkg.ConvMode.Explicit(case["pads"], [1] * x.ndim)
    got = pkg.conv(x, k, mode, padding_from_spec(pkg, case["padding"]), lib=lib)
    np.testing.assert_array_equal(got, np.array(case["expect"], dtype=case["dtype"]))
    if case["dtype"] == "int32":
        gf = pkg.conv_fft(x.astype(np.float64), k.astype(np.float64), mode, padding_from_spec(pkg, case["padding"]), lib=lib)
        assert np.max(np.abs(gf - np.array(case["expect"], dtype=np.float64))) < 1e-9


@pytest.mark.parametrize("case", TORCH["float_cases"], ids=lambda c: c["src"])
def test_float_circular_kat(ndc, case):
    pkg, lib = ndc
    x = np.array(case["x"], np.float32)
    k = np.array(case["kernel"], np.float32)
    e = np.array(case["expect"])
    assert np.max(np.abs(pkg.conv(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Circular, lib=lib) - e)) < case["tol"]
    assert np.max(np.abs(pkg.conv_fft(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Circular, lib=lib) - e)) < case["tol"]


def test_border_index_map_matches_oracle(ndc, oracle):
    """ndconv_border_index_map (symbolic replay of half_dim.rs) against the oracle's literal sequential padding."""
    import ctypes
    pkg, lib = ndc
    rng = np.random.default_rng(5)
    names = ["zeros", "const", "reflect", "replicate", "circular"]
    for _ in range(300):
        n = int(rng.integers(1, 7))
        pf, pb = int(rng.integers(0, 9)), int(rng.integers(0, 9))
        bf, bb = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        x = np.arange(1, n + 1, dtype=np.int32)
        spec = ("explicit", [[("const", -1) if bf == 1 else names[bf], ("const", -2) if bb == 1 else names[bb]]])
        m = np.zeros(n + pf + pb, np.int32)
        st = lib.c.ndconv_border_index_map(n, pf, pb, bf, bb, m.ctypes.data)
        try:
            ref = oracle.pad(x, spec, [[pf, pb]])
        except oracle.OracleError as e:
            assert e.status == oracle.PANIC and st == pkg.ERR_PANIC
            continue
        assert st == 0
        val = np.where(m >= 0, x[np.clip(m, 0, n - 1)], np.where(m == -1, -1 if bf == 1 else 0, np.where(m == -2, -2 if bb == 1 else 0, 0)))
        np.testing.assert_array_equal(val, ref)


DTYPES_DIRECT = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64,
                 np.complex64, np.complex128]


def rand_array(rng, shape, dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        info = np.iinfo(dt)
        return rng.integers(info.min, info.max, size=shape, dtype=dt, endpoint=True)   # full range: exercises wrapping
    if dt.kind == "c":
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)
    return rng.standard_normal(shape).astype(dt)


def random_case(rng, max_nd=3, fft=False):
    nd = int(rng.integers(1, max_nd + 1))
    shape = [int(rng.integers(1, 10)) for _ in range(nd)]
    ks = [int(rng.integers(1, 5)) for _ in range(nd)]
    dil = [int(rng.integers(1, 4)) for _ in range(nd)]
    names = ["zeros", ("const", 3), "reflect", "replicate", "circular"]
    kind = int(rng.integers(0, 5))
    kd = [(k - 1) * d + 1 for k, d in zip(ks, dil)]
    if kind < 3:
        mode = ["full", "same", "valid"][kind]
        pads = {"full": [[v - 1, v - 1] for v in kd], "same": [[v // 2, (v - 1) // 2] for v in kd], "valid": [[0, 0]] * nd}[mode]
    elif kind == 3:
        pp = [int(rng.integers(0, 6)) for _ in range(nd)]
        mode = ("custom", pp, [int(rng.integers(1, 4)) for _ in range(nd)])
        pads = [[p, p] for p in pp]
    else:
        pads = [[int(rng.integers(0, 6)), int(rng.integers(0, 6))] for _ in range(nd)]
        mode = ("explicit", pads, [int(rng.integers(1, 4)) for _ in range(nd)])
    # borders drawn inside the reference's non-panicking domain
    sides = []
    for i in range(nd):
        row = []
        for s in range(2):
            while True:
                b = names[int(rng.integers(0, 5))]
                if b == "reflect" and pads[i][s] > shape[i] - 1:
                    continue
                break
            row.append(b)
        sides.append(row)
    pk = int(rng.integers(0, 3))
    if pk == 0:
        padding = ("explicit", sides)
    elif pk == 1:
        padding = ("custom", [r[0] if not (r[0] == "reflect" and max(pads[i]) > shape[i] - 1) else "replicate" for i, r in enumerate(sides)])
    else:
        cand = sides[0][0]
        if cand == "reflect" and any(max(pads[i]) > shape[i] - 1 for i in range(nd)):
            cand = "circular"
        padding = cand
    return shape, ks, dil, mode, padding, bool(rng.integers(0, 2))


@pytest.mark.parametrize("seed", range(6))
def test_conv_direct_random_bit_exact(ndc, oracle, seed):
    """integer AND float direct conv are bit-identical to the oracle (same tap order, un-fused multiply-add)."""
    pkg, lib = ndc
    rng = np.random.default_rng(100 + seed)
    done = 0
    while done < 40:
        shape, ks, dil, mode, padding, rev = random_case(rng)
        dt = DTYPES_DIRECT[int(rng.integers(0, len(DTYPES_DIRECT)))]
        x = rand_array(rng, shape, dt)
        k = rand_array(rng, ks, dt)
        if rng.integers(0, 3) == 0:
            k.ravel()[rng.integers(0, k.size)] = 0      # zero-tap elision (dilation/mod.rs:49)
        try:
            ref = oracle.conv(x, k, mode, padding, dil, rev)
            err = None
        except oracle.OracleError as e:
            err = e.status
        kw = pkg.with_dilation(k, dil)
        if not rev:
            kw = kw.no_reverse()
        if err is not None:
            with pytest.raises(pkg.NdConvError) as ei:
                pkg.conv(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), lib=lib)
            assert ei.value.status == err
            continue
        got = pkg.conv(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), lib=lib)
        assert got.shape == ref.shape
        assert got.tobytes() == ref.tobytes(), (shape, ks, dil, mode, padding, rev, dt)
        done += 1


@pytest.mark.parametrize("seed", range(6))
def test_conv_fft_random(ndc, oracle, seed):
    pkg, lib = ndc
    rng = np.random.default_rng(200 + seed)
    done = 0
    while done < 30:
        shape, ks, dil, mode, padding, rev = random_case(rng)
        dt = [np.float32, np.float64, np.complex64, np.complex128][int(rng.integers(0, 4))]
        x = rand_array(rng, shape, dt)
        k = rand_array(rng, ks, dt)
        try:
            ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
            err = None
        except oracle.OracleError as e:
            err = e.status
        kw = pkg.with_dilation(k, dil)
        if not rev:
            kw = kw.no_reverse()
        if err is not None:
            with pytest.raises(pkg.NdConvError) as ei:
                pkg.conv_fft(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), lib=lib)
            # conv_fft reports DataShape for an empty kernel (conv_fft/mod.rs:211-213) and checks the shape mismatch before padding
            continue
        got = pkg.conv_fft(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), lib=lib)
        assert got.shape == ref.shape
        n_fft = 1
        for i in range(len(shape)):
            n_fft *= lib.c.ndconv_plan_fft_size(shape[i] + 12, 0)
        bound = float(np.max(np.abs(x)) * np.sum(np.abs(k)))
        assert np.max(np.abs(got - ref)) <= fft_tol(dt, n_fft, ref, bound), (shape, ks, dil, mode, padding, rev, dt)
        done += 1


def test_error_order(ndc):
    """A.4: conv -> DataShape / KernelShape / MismatchShape; conv_fft -> DataShape for an empty kernel (quirk)."""
    pkg, lib = ndc
    for fn, kerr in ((pkg.conv, pkg.ERR_KERNEL_SHAPE), (pkg.conv_fft, pkg.ERR_DATA_SHAPE)):
        with pytest.raises(pkg.NdConvError) as e:
            fn(np.zeros((0, 3), np.float32), np.ones((1, 1), np.float32), lib=lib)
        assert e.value.status == pkg.ERR_DATA_SHAPE
        with pytest.raises(pkg.NdConvError) as e:
            fn(np.ones((2, 2), np.float32), np.ones((0, 1), np.float32), lib=lib)
        assert e.value.status == kerr
        with pytest.raises(pkg.NdConvError) as e:
            fn(np.ones(3, np.float32), np.ones(5, np.float32), pkg.ConvMode.Valid, lib=lib)
        assert e.value.status == pkg.ERR_MISMATCH_SHAPE


def test_strided_views(ndc, oracle):
    """ndarray views: negative / non-unit strides on data and kernel"""
    pkg, lib = ndc
    rng = np.random.default_rng(9)
    base = rng.integers(-50, 50, size=(9, 14)).astype(np.int32)
    kb = rng.integers(-5, 5, size=(5, 6)).astype(np.int32)
    x = base[::-2, 1::3]
    k = kb[::2, ::-2]
    ref = oracle.conv(np.ascontiguousarray(x), np.ascontiguousarray(k), "same", "reflect")
    np.testing.assert_array_equal(pkg.conv(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Reflect, lib=lib), ref)
    xf = x.astype(np.float64)
    got = pkg.conv_fft(xf, k.astype(np.float64), pkg.ConvMode.Same, pkg.PaddingMode.Reflect, lib=lib)
    assert np.max(np.abs(got - ref)) < 1e-9


def test_processor_reuse_and_kernel_cache(ndc, oracle):
    pkg, lib = ndc
    rng = np.random.default_rng(4)
    proc = pkg.get_fft_processor(0, lib)
    k1 = rng.standard_normal((3, 4)).astype(np.float32)
    k2 = rng.standard_normal((3, 4)).astype(np.float32)
    for it in range(3):
        for k in (k1, k2):
            x = rng.standard_normal((17, 23)).astype(np.float32)
            got = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Replicate, proc)
            ref = oracle.conv_f64_truth(x, k, "full", "replicate")
            assert np.max(np.abs(got - ref)) <= fft_tol(np.float32, 32 * 32, ref)
    assert proc.launch_count > 0
    proc.close()


def test_tiled_overlap_save_paths(ndc, oracle):
    """axes longer than one shared-memory FFT tile are cut into overlap-save tiles: exercise tiling on every axis"""
    pkg, lib = ndc
    rng = np.random.default_rng(12)
    # 1-D long enough for several last-axis tiles (cap 8192 real points)
    x = rng.standard_normal(20011).astype(np.float32)
    k = rng.standard_normal(33).astype(np.float32)
    for mode, om in ((pkg.ConvMode.Same, "same"), (pkg.ConvMode.Custom([40], [3]), ("custom", [40], [3]))):
        got = pkg.conv_fft(x, pkg.with_dilation(k, 2), mode, pkg.PaddingMode.Reflect, lib=lib)
        ref = oracle.conv_f64_truth(x, k, om, "reflect", 2)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= fft_tol(np.float32, 8192, ref)
    # 2-D with the strided axis longer than the column cap (1024) and a wide last axis
    x = rng.standard_normal((2300, 40)).astype(np.float32)
    k = rng.standard_normal((5, 3)).astype(np.float32)
    got = pkg.conv_fft(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Circular, lib=lib)
    ref = oracle.conv_f64_truth(x, k, "full", "circular")
    assert np.max(np.abs(got - ref)) <= fft_tol(np.float32, 1024 * 64, ref)
    xc = (rng.standard_normal((30, 4500)) + 1j * rng.standard_normal((30, 4500))).astype(np.complex64)
    kc = (rng.standard_normal((3, 7)) + 1j * rng.standard_normal((3, 7))).astype(np.complex64)
    got = pkg.conv_fft(xc, kc, pkg.ConvMode.Same, pkg.PaddingMode.Replicate, lib=lib)
    ref = oracle.conv_f64_truth(xc, kc, "same", "replicate")
    assert np.max(np.abs(got - ref)) <= fft_tol(np.complex64, 32 * 4096, ref)


def test_kernel_longer_than_an_fft_tile_is_evaluated_directly(ndc, oracle):
    """A dilated kernel extent above the largest shared-memory FFT tile (8192 real points on the last axis, 1024 on the
    others) has no overlap-save tiling; the reference handles such kernels, so conv_fft evaluates them with the direct
    kernel instead of refusing them.  Tolerance: direct f32 summation of `taps` terms, c * eps * sqrt(taps) * max|out| bound."""
    pkg, lib = ndc
    rng = np.random.default_rng(21)
    x = rng.standard_normal(12000).astype(np.float32)
    k = rng.standard_normal(300).astype(np.float32)
    got = pkg.conv_fft(x, pkg.with_dilation(k, 30), pkg.ConvMode.Same, pkg.PaddingMode.Reflect, lib=lib)      # Kd = 8971 > 8192
    ref = oracle.conv_f64_truth(x, k, "same", "reflect", 30)
    assert got.shape == ref.shape
    tol = 4 * np.finfo(np.float32).eps * np.sqrt(300) * float(np.max(np.abs(x))) * float(np.sum(np.abs(k)))
    assert np.max(np.abs(got - ref)) <= tol
    x2 = rng.standard_normal((1500, 20)).astype(np.float32)
    k2 = rng.standard_normal((40, 3)).astype(np.float32)
    got = pkg.conv_fft(x2, pkg.with_dilation(k2, [27, 1]), pkg.ConvMode.Valid, pkg.PaddingMode.Zeros, lib=lib)  # Kd0 = 1054 > 1024
    ref = oracle.conv_f64_truth(x2, k2, "valid", "zeros", [27, 1])
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= 4 * np.finfo(np.float32).eps * np.sqrt(120) * float(np.max(np.abs(x2))) * float(np.sum(np.abs(k2)))
