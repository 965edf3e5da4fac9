"""N > 1 path on CPU: world_size-2 gloo ranks each convolve their overlap-save slab (kernel bodies under host emulation --
test infrastructure) and the gathered rows must equal the single-rank result / the oracle: bit-exact for integers,
within tolerance for floats."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CASES = [
    # dtype, shape, kernel shape, dilation, mode, padding, path
    ("int32", (23, 9), (3, 2), 1, ("custom", [2, 1], [2, 1]), "reflect", "direct"),
    ("int64", (17, 5, 4), (4, 2, 2), 2, "full", ("custom", ["circular", ("const", 7), "replicate"]), "direct"),
    ("float64", (31, 12), (5, 3), 1, "same", ("explicit", [[("const", 1.5), "replicate"], ["reflect", "zeros"]]), "fft"),
    ("float32", (40,), (7,), 1, "full", "circular", "fft"),
]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    sys.path.insert(0, str(ROOT / "tests" / "emul"))
    import ctypes
    import torch
    import torch.distributed as dist
    import build_emul
    from test_parity_small import mode_from_spec, padding_from_spec
    pkg = importlib.import_module("ndarray-conv_b200")
    sharded = importlib.import_module("ndarray-conv_b200.sharded")
    lib = pkg.Library(ctypes.CDLL(str(build_emul.build())))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    results = []
    for ci, (dt, shape, ks, dil, mode, padding, path) in enumerate(CASES):
        rng = np.random.default_rng(50 + ci)          # same data on every rank (host-resident input is cut locally)
        x = rng.integers(-20, 20, size=shape).astype(dt)
        k = rng.integers(-4, 5, size=ks).astype(dt)
        kw = pkg.with_dilation(k, dil)
        y, ob, oe = sharded.conv_rank(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), rank, world,
                                      pkg.PATH_DIRECT if path == "direct" else pkg.PATH_FFT, lib=lib)
        gathered = [None] * world
        dist.all_gather_object(gathered, (ob, oe, y))
        full = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0]) if g[2] is not None], axis=0)
        results.append(full)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put(results)


def test_two_rank_slabs_match_oracle(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for ci, (dt, shape, ks, dil, mode, padding, path) in enumerate(CASES):
        rng = np.random.default_rng(50 + ci)
        x = rng.integers(-20, 20, size=shape).astype(dt)
        k = rng.integers(-4, 5, size=ks).astype(dt)
        ref = oracle.conv(x, k, mode, padding, dil, True)
        got = results[ci]
        assert got.shape == ref.shape
        if path == "direct":
            assert np.array_equal(got, ref)
        else:
            assert np.max(np.abs(got - ref)) <= 1e-4 * max(1.0, float(np.max(np.abs(ref))))


def test_slab_plan_covers_all_rows(pkg):
    lib = pkg.get_library()
    k = np.ones((5, 3), np.float32)
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            sl = pkg.slab_plan((100, 40), np.float32, k, pkg.ConvMode.Custom([3, 1], [2, 1]), pkg.PaddingMode.Reflect, pkg.PATH_FFT, world, r, lib)
            seen += list(range(sl["out_begin"], sl["out_end"]))
            if sl["out_end"] > sl["out_begin"]:
                assert sl["pad_begin"] == sl["out_begin"] * 2 and sl["pad_end"] == (sl["out_end"] - 1) * 2 + 5
        assert seen == list(range((100 + 6 - 5) // 2 + 1))


def _halo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from test_parity_small import mode_from_spec, padding_from_spec
    pkg = importlib.import_module("ndarray-conv_b200")
    sharded = importlib.import_module("ndarray-conv_b200.sharded")
    lib = pkg.get_library()                       # host-side planning only (no compute call)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for padding in ("reflect", "circular", ("explicit", [[("const", 3.0), "replicate"], ["zeros", "zeros"]])):
        n0 = 29
        x = np.random.default_rng(5).random((n0, 6)).astype(np.float32)
        k = np.ones((7, 3), np.float32)
        mode, pm = pkg.ConvMode.Full, padding_from_spec(pkg, padding)
        plans = [sharded.plan_rank(x.shape, x.dtype, k, mode, pm, pkg.PATH_FFT, r, world, lib) for r in range(world)]
        b = sharded.row_partition(n0, world)
        x_local = torch.from_numpy(x[b[rank]:b[rank + 1]].copy())
        slab = sharded.exchange_halo_rows(x_local, n0, plans, rank, world).numpy()
        ok = ok and np.array_equal(slab, sharded.cut_slab(x, plans[rank]))
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    dist.destroy_process_group()
    if rank == 0:
        q.put(all(flags))


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_assembles_the_same_slab(world):
    """device-resident mode: rows are partitioned over ranks; one batched send/recv must rebuild exactly the slab that
    cutting the full host array would give (Reflect / Circular wrap / Const borders)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() + world) % 90
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    assert q.get(timeout=180) is True
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
