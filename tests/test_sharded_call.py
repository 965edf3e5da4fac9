"""ndconv_conv_fft_sharded: conv_fft_par with more than one GPU configured -- one host-resident convolution,
one axis-0 slab of output rows per processor handle, each handle on its own host thread, no data-path collective.
Parity: the slabs tile the output exactly, so the result must equal the single-processor conv_fft up to rounding (a slab may
pick other overlap-save tile lengths than the whole array), (the single-processor path itself is pinned to the oracle by test_parity_*.py and test_baseline_configs.py)."""
import ctypes

import numpy as np
import pytest

from test_parity_small import fft_tol


def test_small_problem_runs_on_first_processor(ndc, oracle):
    """too small to pipeline: the call is ndconv_conv_fft on processors[0]"""
    pkg, lib = ndc
    rng = np.random.default_rng(5)
    x, k = rng.random((40, 50), dtype=np.float32), rng.random((5, 7), dtype=np.float32)
    procs = [pkg.get_fft_processor(0, lib) for _ in range(3)]
    got = pkg.conv_fft_sharded(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs)
    one = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs[0])
    assert np.array_equal(got, one)
    ref = oracle.conv_f64_truth(x, k, "full", "reflect")
    assert np.max(np.abs(got - ref)) <= fft_tol(np.float32, 64 * 64, ref)
    assert procs[1].launch_count == 0 and procs[2].launch_count == 0
    for p in procs:
        p.close()


def test_argument_errors(ndc):
    pkg, lib = ndc
    x, k = np.ones((8, 8), np.float32), np.ones((3, 3), np.float32)
    pr, keep = pkg.make_problem(x.shape, [8, 1], x.ctypes.data, x.dtype, pkg._into_kwd(k), pkg.ConvMode.Same, pkg.PaddingMode.Zeros, pkg.MEM_HOST, lib)
    out = np.empty((8, 8), np.float32)
    assert lib.c.ndconv_conv_fft_sharded(None, 0, ctypes.byref(pr), out.ctypes.data) == pkg.ERR_BAD_ARG
    handles = (ctypes.c_void_p * 2)(None, None)
    assert lib.c.ndconv_conv_fft_sharded(handles, 2, ctypes.byref(pr), out.ctypes.data) == pkg.ERR_BAD_ARG
    # shape errors surface as in conv_fft: Kd > P -> MismatchShape (conv_fft/mod.rs:222-227)
    proc = pkg.get_fft_processor(0, lib)
    with pytest.raises(pkg.NdConvError) as e:
        pkg.conv_fft_sharded(np.ones((2, 2), np.float32), np.ones((3, 3), np.float32), pkg.ConvMode.Valid, pkg.PaddingMode.Zeros, [proc, proc])
    assert e.value.status == pkg.ERR_MISMATCH_SHAPE
    proc.close()


CASES = [
    # shape, kernel, mode, padding, dilation
    ((6000, 5000), (31, 31), "Full", "Reflect", 1),
    ((5000, 6000), (11, 31), "Same", ("Custom", ["Circular", "Replicate"]), 2),
    ((7001, 4099), (63, 5), "Valid", "Zeros", 1),
    ((300, 400, 500), (5, 7, 9), "Same", ("Const", 0.5), 1),
]


def _modes(pkg, mode, padding):
    cm = getattr(pkg.ConvMode, mode)
    if isinstance(padding, tuple) and padding[0] == "Custom":
        pm = pkg.PaddingMode.Custom([getattr(pkg.BorderType, b) for b in padding[1]])
    elif isinstance(padding, tuple):
        pm = pkg.PaddingMode.Const(padding[1])
    else:
        pm = getattr(pkg.PaddingMode, padding)
    return cm, pm


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 3])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c[0])) + "-" + c[2])
def test_sharded_equals_single(pkg, cuda_lib, case, nproc):
    """every handle on every visible device in turn (one GPU: the handles share device 0 and the slabs interleave on its streams)"""
    shape, kshape, mode, padding, dil = case
    rng = np.random.default_rng(6)
    x, k = rng.random(shape, dtype=np.float32), rng.random(kshape, dtype=np.float32)
    cm, pm = _modes(pkg, mode, padding)
    kern = pkg.with_dilation(k, dil) if dil > 1 else k
    ndev = cuda_lib.c.ndconv_device_count()
    print(f"[sharded_call] visible GPUs: {ndev} -> {nproc} handles on {'distinct devices' if ndev >= nproc else ('devices ' + str([i % ndev for i in range(nproc)]))}")
    procs = [pkg.get_fft_processor(i % ndev, cuda_lib) for i in range(nproc)]
    got = pkg.conv_fft_sharded(x, kern, cm, pm, procs)
    assert all(p.launch_count > 0 for p in procs)           # every handle produced its rows
    single = pkg.get_fft_processor(0, cuda_lib)
    one = pkg.conv_fft_with_processor(x, kern, cm, pm, single)
    assert got.shape == one.shape
    tol = fft_tol(np.float32, 1024 * 2048, one)
    assert np.max(np.abs(got - one)) <= tol
    for p in procs + [single]:
        p.close()


@pytest.mark.gpu
def test_sharded_pinned_buffers_and_all_devices(pkg, cuda_lib):
    """pinned host input / output (ndconv_host_alloc) and one handle per visible device: the multi-GPU form of the call"""
    ndev = cuda_lib.c.ndconv_device_count()
    n0, n1, kk = 8192, 4096, 31
    nbytes_in, nbytes_out = n0 * n1 * 4, (n0 + kk - 1) * (n1 + kk - 1) * 4
    pin, pout = cuda_lib.c.ndconv_host_alloc(nbytes_in), cuda_lib.c.ndconv_host_alloc(nbytes_out)
    assert pin and pout
    try:
        x = np.frombuffer((ctypes.c_char * nbytes_in).from_address(pin), dtype=np.float32).reshape(n0, n1)
        out = np.frombuffer((ctypes.c_char * nbytes_out).from_address(pout), dtype=np.float32).reshape(n0 + kk - 1, n1 + kk - 1)
        rng = np.random.default_rng(7)
        x[:] = rng.random((n0, n1), dtype=np.float32)
        k = rng.random((kk, kk), dtype=np.float32)
        procs = [pkg.get_fft_processor(d, cuda_lib) for d in range(max(ndev, 2))] if ndev > 1 else [pkg.get_fft_processor(0, cuda_lib) for _ in range(2)]
        got = pkg.conv_fft_sharded(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs, out=out)
        one = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs[0])
        assert np.max(np.abs(got - one)) <= fft_tol(np.float32, 1024 * 2048, one)
        # constant input under Reflect: every output is c * sum(k)
        x[:] = 0.25
        got = pkg.conv_fft_sharded(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, procs, out=out)
        assert np.max(np.abs(got - 0.25 * k.astype(np.float64).sum())) <= fft_tol(np.float32, 1024 * 2048, got)
        for p in procs:
            p.close()
    finally:
        cuda_lib.c.ndconv_host_free(ctypes.c_void_p(pin))
        cuda_lib.c.ndconv_host_free(ctypes.c_void_p(pout))


def test_batch_distributed_whole(ndc, oracle):
    """ndconv_conv_fft_batch: independent problems (different shapes) round-robin over three handles, one host thread each"""
    pkg, lib = ndc
    rng = np.random.default_rng(8)
    shapes = [(40, 50), (33, 70), (64, 64), (20, 90), (51, 37), (40, 50), (8, 200)]
    xs = [rng.random(s, dtype=np.float32) for s in shapes]
    ks = [rng.random((3 + i % 3, 5), dtype=np.float32) for i in range(len(xs))]
    procs = [pkg.get_fft_processor(0, lib) for _ in range(3)]
    outs = pkg.conv_fft_batch(xs, ks, pkg.ConvMode.Same, pkg.PaddingMode.Replicate, procs)
    assert all(p.launch_count > 0 for p in procs)
    for x, k, y in zip(xs, ks, outs):
        ref = oracle.conv_f64_truth(x, k, "same", "replicate")
        assert y.shape == ref.shape and np.max(np.abs(y - ref)) <= fft_tol(np.float32, 128 * 256, ref)
    # one kernel for all, fewer problems than handles; a failing problem reports its status
    outs = pkg.conv_fft_batch(xs[:2], ks[0], pkg.ConvMode.Full, pkg.PaddingMode.Zeros, procs)
    assert outs[0].shape == (40 + ks[0].shape[0] - 1, 54)
    with pytest.raises(pkg.NdConvError) as e:
        pkg.conv_fft_batch([xs[0], np.ones((2, 2), np.float32)], np.ones((3, 3), np.float32), pkg.ConvMode.Valid, pkg.PaddingMode.Zeros, procs)
    assert e.value.status == pkg.ERR_MISMATCH_SHAPE
    for p in procs:
        p.close()


@pytest.mark.gpu
def test_registered_host_arrays(pkg, cuda_lib):
    """ndconv_host_register: an ordinary numpy array page-locked in place gives the same result as the pageable call"""
    rng = np.random.default_rng(9)
    x, k = rng.random((4000, 7000), dtype=np.float32), rng.random((9, 9), dtype=np.float32)
    proc = pkg.get_fft_processor(0, cuda_lib)
    a = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Reflect, proc)
    with pkg.pinned(x, lib=cuda_lib):
        b = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Reflect, proc)
    assert np.array_equal(a, b)
    assert cuda_lib.c.ndconv_host_register(None, 16) == pkg.ERR_BAD_ARG
    proc.close()


def test_pinned_context_manager_under_emulation(pkg, emul_lib):
    """pkg.pinned(): register / unregister around host calls (a no-op under host emulation; the device build's cudaHostRegister is
    covered by test_registered_host_arrays)"""
    lib = emul_lib
    assert lib.c.ndconv_host_register(None, 16) == pkg.ERR_BAD_ARG
    assert lib.c.ndconv_host_unregister(None) == pkg.ERR_BAD_ARG
    x, k = np.arange(48, dtype=np.float32).reshape(6, 8), np.ones((3, 3), np.float32)
    a = pkg.conv_fft(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Zeros, lib=lib)
    with pkg.pinned(x, lib=lib):
        b = pkg.conv_fft(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Zeros, lib=lib)
    assert np.array_equal(a, b)
