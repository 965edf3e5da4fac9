"""The overlap-save planner through ndconv_plan_query (host logic only: runs on the GPU-less box against the product library).
Any tile length F >= Kd is a valid FFT size (the reference's good_size choice is unobservable), so what is pinned here is
THIS build's choice for the BASELINE configurations -- the tiles DESIGN.md section 3.3 / 5 quote and the profiles were taken with."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def q(pkg, cuda_lib):
    return lambda *a, **kw: pkg.plan_query(*a, lib=cuda_lib, **kw)


def z(shape, dt=np.float32):
    return np.zeros(shape, dt)


def test_c5_tiles_workspace_and_tail_split(pkg, q):
    p = q((32768, 32768), np.float32, z((63, 63)), pkg.ConvMode.Full, pkg.PaddingMode.Reflect)
    assert p["path"] == "fast" and p["tile_len"] == [1024, 2048] and p["tile_valid"] == [962, 1986] and p["n_tiles"] == [35, 17]
    assert p["workspace_bytes"] == 35 * 17 * 1024 * (1024 + 8) * 8          # 5.03 GB: rows of L + 8 complex
    assert p["split_out_rows"] == 34 * 962                                   # 32830 output rows = 34 exact tile rows + a 122-row tail
    assert not p["pipelined"]
    assert q((32768, 32768), np.float32, z((63, 63)), pkg.ConvMode.Full, pkg.PaddingMode.Reflect, memory=pkg.MEM_HOST)["pipelined"]


def test_one_rank_of_c5_on_8_gpus(pkg, q):
    """4104 output rows + halo: 4 full tile rows and a tail that gets a 512-row tile of its own (axis-0 split)"""
    p = q((4166, 32768), np.float32, z((63, 63)), pkg.ConvMode.Explicit([[0, 0], [62, 62]], [1, 1]), pkg.PaddingMode.Reflect)
    assert p["tile_len"] == [1024, 2048] and p["n_tiles"] == [5, 17] and p["split_out_rows"] == 4 * 962


def test_small_problems_take_short_tiles(pkg, q):
    B = pkg.BorderType
    c2 = q((200, 5000), np.float32, pkg.with_dilation(z((11, 31)), 2), pkg.ConvMode.Same, pkg.PaddingMode.Custom([B.Reflect, B.Circular]))
    assert c2["path"] == "fast" and c2["tile_len"] == [128, 256] and c2["n_tiles"] == [2, 26]        # fewest samples would be 256 x 2048
    c3 = q((10, 100, 200), np.float32, z((5, 11, 31)), pkg.ConvMode.Same, pkg.PaddingMode.Zeros)
    assert c3["path"] == "fast" and c3["tile_len"] == [16, 64, 256] and c3["n_tiles"] == [1, 2, 1]
    c3c = q((10, 100, 200), np.complex64, z((5, 11, 31), np.complex64), pkg.ConvMode.Same, pkg.PaddingMode.Zeros)
    assert c3c["tile_len"] == c3["tile_len"] and c3c["workspace_bytes"] == 16 * 128 * 256 * 8          # complex rows: exactly L columns
    c1 = q((5000,), np.float32, z((31,)), pkg.ConvMode.Same, pkg.PaddingMode.Zeros)
    assert c1["path"] == "fast" and c1["tile_len"] == [256] and c1["workspace_bytes"] == 0             # rank 1: one fused launch, no workspace


def test_tile_identities(pkg, q):
    """V = F - Kd + 1 and the tiles cover the alias-free positions: n_tiles * V >= P - Kd + 1 (conv_fft/mod.rs:229-231,282-289)"""
    rng = np.random.default_rng(0)
    for _ in range(60):
        nd = int(rng.integers(1, 4))
        shape = [int(rng.integers(20, 700)) for _ in range(nd)]
        ks = [int(rng.integers(1, 12)) for _ in range(nd)]
        dil = int(rng.integers(1, 3))
        dt = [np.float32, np.float64, np.complex64][int(rng.integers(0, 3))]
        p = q(tuple(shape), dt, pkg.with_dilation(z(ks, dt), dil), pkg.ConvMode.Full, pkg.PaddingMode.Replicate)
        for a in range(nd):
            Kd = (ks[a] - 1) * dil + 1
            P = shape[a] + 2 * (Kd - 1)
            assert p["tile_valid"][a] == p["tile_len"][a] - Kd + 1 >= 1
            assert p["n_tiles"][a] * p["tile_valid"][a] >= P - Kd + 1
            assert (p["n_tiles"][a] - 1) * p["tile_valid"][a] < P - Kd + 1          # no tile without a single useful position


def test_path_selection(pkg, q):
    f64 = q((300, 500), np.float64, z((5, 7), np.float64), pkg.ConvMode.Same, pkg.PaddingMode.Zeros)
    assert f64["path"] == "fast" and f64["tile_len"][1] <= 512 and f64["tile_len"][0] <= 256                                      # f64 fast path: sixteen values per thread
    assert q((300, 500), np.complex128, z((5, 7), np.complex128), pkg.ConvMode.Same, pkg.PaddingMode.Zeros)["path"] == "generic"  # Complex<f64>: generic kernels
    assert q((3, 4, 5, 6), np.float32, z((2, 2, 2, 2)), pkg.ConvMode.Same, pkg.PaddingMode.Zeros)["path"] == "generic"           # rank 4
    long1 = q((100000,), np.float32, z((9001,)), pkg.ConvMode.Valid, pkg.PaddingMode.Zeros)                                      # kernel longer than one FFT tile:
    assert long1["path"] == "split" and long1["n_tiles"] == [3] and long1["tile_valid"] == [4096]                                # cut into segments of <= cap / 2 taps
    assert q((100000,), np.float32, z((9001,)), pkg.ConvMode.Same, pkg.PaddingMode.Circular)["path"] == "direct"                 # a Circular border on the cut axis: direct kernel
    with pytest.raises(pkg.NdConvError) as e:
        q((3, 3), np.float32, z((5, 5)), pkg.ConvMode.Valid, pkg.PaddingMode.Zeros)
    assert e.value.status == pkg.ERR_MISMATCH_SHAPE
