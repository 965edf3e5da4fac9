"""sm_100a fast path (kernels_fft_fast.cuh: real f32 / Complex<f32>, rank 1-3, power-of-two overlap-save tiles) against the CPU oracle.
Device-only kernels (warp shuffles): these tests need a B200."""
import numpy as np
import pytest

from test_parity_small import fft_tol, mode_from_spec, padding_from_spec

pytestmark = pytest.mark.gpu

CASES = [
    # shape, kernel, dilation, mode, padding, reverse
    ((700, 1500), (5, 9), 1, "same", "zeros", True),
    ((1300, 2600), (5, 9), 1, "full", "reflect", True),
    ((1300, 2600), (4, 6), 2, "same", ("custom", ["circular", ("const", 1.5)]), False),
    ((2100, 4200), (7, 3), 1, ("custom", [3, 5], [2, 3]), "replicate", True),
    ((1025, 2049), (63, 63), 1, "valid", "zeros", True),
    ((900, 5000), (3, 31), 3, ("explicit", [[0, 7], [40, 2]], [1, 1]), ("explicit", [["zeros", "reflect"], ["circular", "replicate"]]), True),
    # 256-row column tiles (radix 16 x 16)
    ((200, 5000), (11, 31), 2, "same", ("custom", ["reflect", "circular"]), True),          # BASELINE configs[1]
    ((130, 1300), (3, 5), 1, "full", "replicate", False),
    ((470, 1500), (9, 4), 1, ("custom", [4, 0], [3, 2]), ("const", -0.5), True),             # two 256-row tiles, strided output
    # small last axes (T = 4, 8, 16 lanes per row) and odd column tiles
    ((70, 150), (3, 3), 1, "same", "reflect", True),
    ((33, 400), (5, 9), 1, "full", ("custom", ["circular", "replicate"]), False),
    ((600, 700), (7, 7), 2, "valid", "zeros", True),
    ((50, 300), (17, 3), 1, "same", ("const", 2.0), True),
    # rank 3
    ((10, 100, 200), (5, 11, 31), 1, "same", "zeros", True),                                 # BASELINE configs[2]
    ((20, 70, 300), (3, 4, 5), 1, "full", ("custom", ["reflect", "circular", ("const", 0.5)]), False),
    ((40, 300, 130), (2, 3, 9), 2, ("custom", [1, 0, 5], [2, 1, 3]), "replicate", True),
    ((300, 20, 520), (9, 3, 3), 1, "same", ("explicit", [["zeros", "reflect"], ["replicate", "circular"], ["reflect", "zeros"]]), True),
    # axis-0 split (plan_axis0_split): the last axis-0 tile would be mostly padding, so the rows are cut into two sub-convolutions
    ((1100, 4000), (5, 9), 1, "same", "reflect", True),
    ((1100, 4000), (5, 9), 1, ("custom", [2, 4], [3, 2]), "replicate", False),                 # strided outputs across the cut
    ((1100, 4000), (5, 9), 2, "full", ("custom", [("const", 0.75), "circular"]), True),
    ((1100, 4000), (5, 9), 1, "same", ("custom", ["circular", "reflect"]), True),             # Circular on axis 0: never split
    ((600, 40, 300), (5, 3, 3), 1, "same", ("const", 0.25), True),                             # rank 3
    # 1024-row column tiles outside the 2-D case (col_pass_tma, modes FWD / INV, outer > 1, row pitch != 1032)
    ((3, 1000, 300), (2, 3, 5), 1, "same", "reflect", True),                                   # middle axis: forward and inverse passes of their own
    ((1000, 20, 300), (3, 3, 3), 1, "full", ("custom", ["replicate", "circular", "zeros"]), False),   # axis 0 with a [F1][pitch] inner extent
    ((1000, 1500), (4, 4), 1, "same", "zeros", True),                                          # even Kd: odd number of aliased head rows
    # rank 1: fast::row1d, the whole pipeline in one launch (tiles of 256 .. 2048 samples)
    ((5000,), (31,), 1, "same", "zeros", True),                                                # BASELINE configs[0]
    ((700,), (5,), 1, "full", "reflect", True),
    ((1000,), (100,), 1, "valid", "zeros", False),
    ((3000,), (7,), 3, ("custom", [9], [2]), "replicate", True),
    ((100000,), (63,), 1, "full", ("const", -1.5), True),
    ((40000,), (33,), 2, "same", "circular", False),
    ((65536,), (1000,), 1, ("explicit", [[17, 400]], [3]), ("explicit", [["reflect", "replicate"]]), True),
]


@pytest.mark.parametrize("case", CASES, ids=[str(c[:2]) for c in CASES])
def test_opt2d_vs_oracle(pkg, cuda_lib, oracle, case):
    shape, ks, dil, mode, padding, rev = case
    rng = np.random.default_rng(42)
    x = rng.random(shape, dtype=np.float32) - 0.25
    k = rng.random(ks, dtype=np.float32) - 0.5
    kw = pkg.with_dilation(k, dil)
    if not rev:
        kw = kw.no_reverse()
    proc = pkg.get_fft_processor(0, cuda_lib)
    got = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
    got2 = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)   # cached kernel spectrum
    ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
    assert got.shape == ref.shape
    tol = fft_tol(np.float32, 1024 * 2048, ref, float(np.max(np.abs(x)) * np.sum(np.abs(k))))
    assert np.max(np.abs(got - ref)) <= tol, (np.max(np.abs(got - ref)), tol)
    assert np.array_equal(got, got2)
    proc.close()


CX_CASES = [
    # Complex<f32>: C2C rows (fast::row_fwd_c / row_inv_c), last-axis tiles of 128 .. 1024 samples
    ((10, 100, 200), (5, 11, 31), 1, "same", "zeros", True),                                 # BASELINE configs[2], Complex variant
    ((300, 700), (5, 9), 1, "full", "reflect", True),
    ((130, 2500), (4, 6), 2, "same", ("custom", ["circular", ("const", 1.5 - 0.75j)]), False),
    ((1100, 1300), (7, 3), 1, ("custom", [3, 5], [2, 3]), "replicate", True),                 # strided; large enough for the axis-0 split
    ((1025, 1030), (63, 63), 1, "valid", ("const", 0.25 + 2j), True),
    ((40, 90, 130), (3, 4, 5), 1, "full", ("custom", ["reflect", ("const", -1j), "circular"]), False),
    ((64, 64), (3, 3), 1, "same", "zeros", True),                                             # below the fast path's size threshold: generic kernels
    # rank 1: fast::row1d_c
    ((5000,), (31,), 1, "same", "zeros", True),
    ((700,), (5,), 2, "full", ("const", 0.5 - 1j), False),
    ((90000,), (400,), 1, ("custom", [7], [3]), "reflect", True),
    ((3000,), (9,), 1, "valid", "circular", True),
]


@pytest.mark.parametrize("case", CX_CASES, ids=[str(c[:2]) for c in CX_CASES])
def test_complex_fast_path_vs_oracle(pkg, cuda_lib, oracle, case):
    shape, ks, dil, mode, padding, rev = case
    rng = np.random.default_rng(7)
    x = ((rng.random(shape) - 0.25) + 1j * (rng.random(shape) - 0.5)).astype(np.complex64)
    k = ((rng.random(ks) - 0.5) + 1j * (rng.random(ks) - 0.25)).astype(np.complex64)
    kw = pkg.with_dilation(k, dil)
    if not rev:
        kw = kw.no_reverse()
    proc = pkg.get_fft_processor(0, cuda_lib)
    got = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
    got2 = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
    ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
    assert got.shape == ref.shape and got.dtype == np.complex64
    tol = fft_tol(np.complex64, 1024 * 1024, ref, float(np.max(np.abs(x)) * np.sum(np.abs(k))))
    assert np.max(np.abs(got - ref)) <= tol, (np.max(np.abs(got - ref)), tol)
    assert np.array_equal(got, got2)
    proc.close()


def test_opt2d_matches_generic_path(pkg, cuda_lib):
    """same input through the fast path and (NDCONV_DISABLE_OPT=1, subprocess) the generic path"""
    import os
    import subprocess
    import sys
    code = (
        "import importlib,numpy as np,sys; sys.path.insert(0,'.');"
        "pkg=importlib.import_module('ndarray-conv_b200');"
        "rng=np.random.default_rng(3); x=rng.random((1200,2500),dtype=np.float32); k=rng.random((9,11),dtype=np.float32);"
        "y=pkg.conv_fft(x,k,pkg.ConvMode.Full,pkg.PaddingMode.Reflect); np.save(sys.argv[1],y)"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tag, env in (("opt", {}), ("gen", {"NDCONV_DISABLE_OPT": "1"})):
        path = f"/tmp/ndconv_{tag}.npy"
        subprocess.run([sys.executable, "-c", code, path], check=True, cwd=root, env={**os.environ, **env})
        outs.append(np.load(path))
    scale = float(np.max(np.abs(outs[1])))
    assert np.max(np.abs(outs[0] - outs[1])) <= 8 * np.finfo(np.float32).eps * np.log2(1024 * 2048) * scale


def test_workspace_budget_stripes(pkg, cuda_lib):
    """a plan whose workspace exceeds the budget (NDCONV_WS_BUDGET_MB, default 24 GB) is processed in stripes of tile rows through
    one workspace: the striped result must equal the unstriped one bit for bit (same tiles, same kernels)"""
    import os
    import subprocess
    import sys
    code = (
        "import importlib,numpy as np,sys; sys.path.insert(0,'.');"
        "pkg=importlib.import_module('ndarray-conv_b200');"
        "rng=np.random.default_rng(5); x=rng.random((3000,2600),dtype=np.float32)-0.5; k=rng.random((7,9),dtype=np.float32);"
        "p=pkg.get_fft_processor(0);"
        "y=pkg.conv_fft_with_processor(x,k,pkg.ConvMode.Custom([5,3],[2,1]),pkg.PaddingMode.Custom([pkg.BorderType.Replicate,pkg.BorderType.Reflect]),p);"
        "np.save(sys.argv[1],y); print('WS', p.workspace_bytes)"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs, ws = [], []
    # NDCONV_TILE_SLACK=-1: a stripe of this (deliberately tiny) array would count as a small problem and pick shorter tiles than
    # the whole; real stripes (workspace over 24 GB) never do
    for tag, env in (("full", {"NDCONV_DISABLE_SPLIT": "1", "NDCONV_TILE_SLACK": "-1"}), ("striped", {"NDCONV_WS_BUDGET_MB": "20", "NDCONV_TILE_SLACK": "-1"})):
        path = f"/tmp/ndconv_ws_{tag}.npy"
        r = subprocess.run([sys.executable, "-c", code, path], check=True, cwd=root, env={**os.environ, **env}, capture_output=True, text=True)
        outs.append(np.load(path))
        ws.append(int(r.stdout.split("WS")[1].split()[0]))
    assert np.array_equal(outs[0], outs[1])
    assert ws[0] - ws[1] >= 20 << 20, ws   # all device buffers of the processor: three tile rows of workspace (3 x 16.9 MB) against one (measured 96.6 vs 69.0 MB)


@pytest.mark.parametrize("padding", ["reflect", ("custom", ["circular", "replicate"]), ("explicit", [[("const", 2.0), "replicate"], ["zeros", "reflect"]])],
                         ids=["reflect", "circular-replicate", "const-mixed"])
def test_pipelined_host_path(pkg, cuda_lib, oracle, padding):
    """host problems >= 96 MB run as overlapped H2D | kernels | D2H slabs; axis-0 borders are materialised while staging"""
    rng = np.random.default_rng(77)
    x = rng.random((6100, 5003), dtype=np.float32)
    k = rng.random((5, 3), dtype=np.float32) - 0.3
    proc = pkg.get_fft_processor(0, cuda_lib)
    got = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Full, padding_from_spec(pkg, padding), proc)
    ref = oracle.conv_f64_truth(x, k, "full", padding)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= fft_tol(np.float32, 1024 * 2048, ref)
    # strided axis-0 output (stride 2) through the same path
    got = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Custom([2, 1], [2, 1]), padding_from_spec(pkg, padding), proc)
    ref = oracle.conv_f64_truth(x, k, ("custom", [2, 1], [2, 1]), padding)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= fft_tol(np.float32, 1024 * 2048, ref)
    proc.close()


def _random_fast_case(rng):
    """a random problem sized for the sm_100a fast path (rank 1-3, f32 or Complex<f32>), borders inside the reference's
    non-panicking domain"""
    nd = int(rng.integers(1, 4))
    if nd == 1:
        shape = [int(rng.integers(600, 200000))]
    elif nd == 2:
        n1 = int(rng.integers(130, 3000))
        shape = [int(rng.integers(20, max(21, min(1500, 2_000_000 // n1)))), n1]
    else:
        n2 = int(rng.integers(130, 600)); n1 = int(rng.integers(16, 200))
        shape = [int(rng.integers(4, max(5, min(60, 2_000_000 // (n1 * n2))))), n1, n2]
    ks = [int(rng.integers(1, 8)) for _ in range(nd)]
    ks[-1] = int(rng.integers(1, 40))
    dil = [int(rng.integers(1, 4)) for _ in range(nd)]
    kd = [(k - 1) * d + 1 for k, d in zip(ks, dil)]
    kind = int(rng.integers(0, 5))
    if kind < 3:
        mode = ["full", "same", "valid"][kind]
        pads = {"full": [[v - 1, v - 1] for v in kd], "same": [[v // 2, (v - 1) // 2] for v in kd], "valid": [[0, 0]] * nd}[mode]
    elif kind == 3:
        pp = [int(rng.integers(0, 12)) for _ in range(nd)]
        mode = ("custom", pp, [int(rng.integers(1, 4)) for _ in range(nd)])
        pads = [[q, q] for q in pp]
    else:
        pads = [[int(rng.integers(0, 12)), int(rng.integers(0, 12))] for _ in range(nd)]
        mode = ("explicit", pads, [int(rng.integers(1, 4)) for _ in range(nd)])
    if any(shape[i] + pads[i][0] + pads[i][1] < kd[i] for i in range(nd)):
        return None
    names = ["zeros", ("const", 1.25), "reflect", "replicate", "circular"]
    sides = []
    for i in range(nd):
        row = []
        for sd in range(2):
            b = names[int(rng.integers(0, 5))]
            if pads[i][sd] > shape[i] - 1 and b in ("reflect", "circular"):
                b = "replicate"
            row.append(b)
        sides.append(row)
    pk = int(rng.integers(0, 3))
    padding = ("explicit", sides) if pk == 0 else (("custom", [r[0] if max(pads[i]) <= shape[i] - 1 or r[0] not in ("reflect", "circular") else "replicate" for i, r in enumerate(sides)]) if pk == 1 else
                                                 (sides[0][0] if all(max(pads[i]) <= shape[i] - 1 for i in range(nd)) or sides[0][0] not in ("reflect", "circular") else "zeros"))
    cx = bool(rng.integers(0, 3) == 0)
    return shape, ks, dil, mode, padding, bool(rng.integers(0, 2)), cx


@pytest.mark.parametrize("seed", range(3))
def test_fast_path_random(pkg, cuda_lib, oracle, seed):
    rng = np.random.default_rng(900 + seed)
    proc = pkg.get_fft_processor(0, cuda_lib)
    done = 0
    while done < 16:
        case = _random_fast_case(rng)
        if case is None:
            continue
        shape, ks, dil, mode, padding, rev, cx = case
        if cx:
            x = ((rng.random(shape) - 0.5) + 1j * (rng.random(shape) - 0.5)).astype(np.complex64)
            k = ((rng.random(ks) - 0.5) + 1j * (rng.random(ks) - 0.5)).astype(np.complex64)
        else:
            x = (rng.random(shape, dtype=np.float32) - 0.5)
            k = (rng.random(ks, dtype=np.float32) - 0.5)
        try:
            ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
        except oracle.OracleError:
            continue
        kw = pkg.with_dilation(k, dil)
        if not rev:
            kw = kw.no_reverse()
        got = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
        assert got.shape == ref.shape, case
        tol = fft_tol(x.dtype, 1024 * 2048, ref, float(np.max(np.abs(x)) * np.sum(np.abs(k))))
        assert np.max(np.abs(got - ref)) <= tol, (case, float(np.max(np.abs(got - ref))), tol)
        done += 1
    proc.close()
