"""sm_100a fast path for real f64 (kernels_fft_fast_f64.cuh: rank 2 / 3, overlap-save tiles of 128 / 256 / 512 samples on the last axis,
16 .. 256 rows on the others) against the CPU oracle.  The reference gates f64 at 1e-9 (conv_fft/tests.rs:15-16); the bound here is
the north star's |err| <= 4 eps log2(N) max|out| with eps = 2^-52, plus the rounding of the direct f64 evaluation it is compared with."""
import numpy as np
import pytest

from test_parity_small import fft_tol, mode_from_spec, padding_from_spec

CASES = [
    # shape, kernel, dilation, mode, padding, reverse
    ((700, 1500), (5, 9), 1, "same", "zeros", True),                                          # 256-row column tiles, 512-sample row tiles
    ((1300, 2600), (5, 9), 1, "full", "reflect", True),
    ((1300, 2600), (4, 6), 2, "same", ("custom", ["circular", ("const", 1.5)]), False),
    ((2100, 1200), (7, 3), 1, ("custom", [3, 5], [2, 3]), "replicate", True),                 # strided outputs
    ((600, 700), (63, 63), 1, "valid", "zeros", True),                                        # the BASELINE kernel extent
    ((900, 5000), (3, 31), 3, ("explicit", [[0, 7], [40, 2]], [1, 1]), ("explicit", [["zeros", "reflect"], ["circular", "replicate"]]), True),
    ((200, 5000), (11, 31), 2, "same", ("custom", ["reflect", "circular"]), True),            # BASELINE configs[1] in f64
    ((130, 1300), (3, 5), 1, "full", "replicate", False),
    ((140, 150), (3, 3), 1, "same", "reflect", True),                                         # 128-sample rows (T = 4), short column tiles
    ((33, 400), (5, 9), 1, "full", ("custom", ["circular", "replicate"]), False),
    ((50, 300), (17, 3), 1, "same", ("const", 2.0), True),
    ((300, 250), (100, 3), 1, "same", "reflect", True),                                       # Kd = 100 on a strided axis: 256-row tiles, 157 useful rows
    # rank 3
    ((10, 100, 200), (5, 11, 31), 1, "same", "zeros", True),                                  # BASELINE configs[2] in f64
    ((20, 70, 300), (3, 4, 5), 1, "full", ("custom", ["reflect", "circular", ("const", 0.5)]), False),
    ((40, 300, 130), (2, 3, 9), 2, ("custom", [1, 0, 5], [2, 1, 3]), "replicate", True),
    ((300, 20, 520), (9, 3, 3), 1, "same", ("explicit", [["zeros", "reflect"], ["replicate", "circular"], ["reflect", "zeros"]]), True),
    # large enough for the axis-0 split (the last axis-0 tile would be mostly padding)
    ((1100, 4000), (5, 9), 1, "same", "reflect", True),
    ((600, 40, 300), (5, 3, 3), 1, "same", ("const", 0.25), True),
]


@pytest.mark.parametrize("case", CASES, ids=[str(c[:2]) for c in CASES])
def test_cases_take_the_f64_fast_path(pkg, case):
    """host logic only: these shapes are planned onto the f64 fast path with tiles from its menus"""
    shape, ks, dil, mode, padding, rev = case
    info = pkg.plan_query(shape, np.float64, pkg.with_dilation(np.ones(ks, np.float64), dil), mode_from_spec(pkg, mode), padding_from_spec(pkg, padding))
    assert info["path"] == "fast", info
    assert info["tile_len"][-1] in (128, 256, 512) and all(f in (16, 32, 64, 128, 256) for f in info["tile_len"][:-1]), info


def test_f64_outside_the_fast_path(pkg):
    z = lambda ks: np.ones(ks, np.float64)
    q = lambda s, k: pkg.plan_query(s, np.float64, k, pkg.ConvMode.Same, pkg.PaddingMode.Zeros)["path"]
    assert q((5000,), z((31,))) == "generic"                     # rank 1
    assert q((400, 2000), z((3, 300))) == "generic"              # Kd > 256 on the last axis
    assert q((2000, 400), z((200, 3))) == "generic"              # Kd > 128 on a strided axis
    assert pkg.plan_query((300, 500), np.complex128, np.ones((5, 7), np.complex128), pkg.ConvMode.Same, pkg.PaddingMode.Zeros)["path"] == "generic"


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[str(c[:2]) for c in CASES])
def test_f64_fast_vs_oracle(pkg, cuda_lib, oracle, case):
    shape, ks, dil, mode, padding, rev = case
    rng = np.random.default_rng(42)
    x = rng.random(shape) - 0.25
    k = rng.random(ks) - 0.5
    kw = pkg.with_dilation(k, dil)
    if not rev:
        kw = kw.no_reverse()
    proc = pkg.get_fft_processor(0, cuda_lib)
    got = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
    got2 = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)   # cached kernel spectrum
    ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
    assert got.shape == ref.shape and got.dtype == np.float64
    bound = float(np.max(np.abs(x)) * np.sum(np.abs(k)))
    tol = fft_tol(np.float64, 512 * 256, ref, bound) + 4 * np.finfo(np.float64).eps * bound          # + the rounding of the direct f64 sum itself
    err = float(np.max(np.abs(got - ref)))
    assert err <= tol, (err, tol)
    assert err <= 1e-9 * max(1.0, float(np.max(np.abs(ref))))                                        # the reference's own gate
    assert np.array_equal(got, got2)
    proc.close()


@pytest.mark.gpu
def test_f64_fast_same_shape_batch(pkg, cuda_lib, oracle):
    """a leading axis of kernel extent 1 is folded into the launches of the f64 fast path as well"""
    rng = np.random.default_rng(3)
    x = rng.random((6, 200, 700)) - 0.5
    k = rng.random((1, 5, 7)) - 0.5
    proc = pkg.get_fft_processor(0, cuda_lib)
    got = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Custom([pkg.BorderType.Zeros, pkg.BorderType.Reflect, pkg.BorderType.Circular]), proc)
    ref = oracle.conv_f64_truth(x, k, "same", ("custom", ["zeros", "reflect", "circular"]), 1, True)
    assert np.max(np.abs(got - ref)) <= 1e-12
    proc.close()
