"""Same-shape batches (the reference's `Baked` sketch, src/conv_fft/mod.rs:22-42): a leading axis of kernel
extent 1 is folded into ONE launch per pass over the tiles of all problems.  The folded call must agree with the problem-by-problem
calls up to rounding (the planner sees the whole stack and may pick longer tiles than for one small problem; with the same tiles --
NDCONV_TILE_SLACK=-1 -- the two are bit-identical, tools/diag_batch.py) and sit within the float gate of the oracle."""
import numpy as np
import pytest

from test_parity_small import fft_tol, mode_from_spec, padding_from_spec

pytestmark = pytest.mark.gpu

CASES = [
    # stack shape, kernel (leading 1s), dilation per axis, mode, padding
    ((6, 200, 5000), (1, 11, 31), (1, 2, 2), "same", ("custom", ["zeros", "reflect", "circular"])),          # BASELINE configs[1] x 6
    ((3, 2, 70, 300), (1, 1, 3, 5), (1, 1, 1, 1), "full", ("custom", ["zeros", "zeros", "replicate", ("const", 0.5)])),   # two folded axes
    ((5, 10, 100, 200), (1, 5, 11, 31), (1, 1, 1, 1), "same", "zeros"),                                        # BASELINE configs[2] x 5 (rank 3 inside)
    ((4, 1100, 1500), (1, 5, 9), (1, 1, 1), ("custom", [0, 2, 4], [1, 3, 2]), "replicate"),                    # 1024-row column tiles, strided outputs
]


@pytest.mark.parametrize("case", CASES, ids=[str(c[0]) for c in CASES])
def test_folded_batch_equals_problem_by_problem(pkg, cuda_lib, oracle, case):
    shape, ks, dil, mode, padding = case
    nb = next(i for i, v in enumerate(ks) if v != 1)
    rng = np.random.default_rng(9)
    x = rng.random(shape, dtype=np.float32) - 0.4
    k = rng.random(ks, dtype=np.float32) - 0.5
    proc = pkg.get_fft_processor(0, cuda_lib)
    cm, pm = mode_from_spec(pkg, mode), padding_from_spec(pkg, padding)
    l0 = proc.launch_count
    got = pkg.conv_fft_with_processor(x, pkg.with_dilation(k, list(dil)), cm, pm, proc)
    folded_launches = proc.launch_count - l0
    # problem by problem with the lower-rank kernel / mode / padding
    def sub_mode(m):
        if isinstance(m, tuple) and m[0] == "custom":
            return ("custom", m[1][nb:], m[2][nb:])
        return m
    def sub_pad(p):
        if isinstance(p, tuple) and p[0] == "custom":
            return ("custom", p[1][nb:])
        return p
    ksub = k.reshape(ks[nb:])
    xs = x.reshape((-1,) + tuple(shape[nb:]))
    cm2, pm2 = mode_from_spec(pkg, sub_mode(mode)), padding_from_spec(pkg, sub_pad(padding))
    l0 = proc.launch_count
    each = [pkg.conv_fft_with_processor(xs[b], pkg.with_dilation(ksub, list(dil[nb:])), cm2, pm2, proc) for b in range(xs.shape[0])]
    each_launches = proc.launch_count - l0
    want = np.stack(each).reshape(got.shape)
    assert folded_launches * 2 <= each_launches, (folded_launches, each_launches)     # one launch per pass for the whole stack (+ the one-off kernel spectrum)
    ref = oracle.conv_f64_truth(xs[1], ksub, sub_mode(mode), sub_pad(padding), list(dil[nb:]) if len(set(dil[nb:])) > 1 else dil[nb], True)
    tol = fft_tol(np.float32, 1024 * 2048, ref, float(np.max(np.abs(x)) * np.sum(np.abs(k))))
    assert np.max(np.abs(got.reshape((-1,) + ref.shape)[1] - ref)) <= tol
    assert np.max(np.abs(got - want)) <= 2 * tol, (float(np.max(np.abs(got - want))), tol)
    proc.close()


def test_folded_batch_device_resident_view(pkg, cuda_lib):
    """device-resident stack with a batch stride larger than one problem (a view into a bigger buffer)"""
    import torch
    rng = np.random.default_rng(4)
    B, n0, n1 = 7, 300, 700
    big = torch.from_numpy(rng.random((B, n0 + 5, n1), dtype=np.float32)).cuda()
    k = rng.random((1, 5, 7), dtype=np.float32)
    proc = pkg.get_fft_processor(0, cuda_lib)
    pm = pkg.PaddingMode.Custom([pkg.BorderType.Zeros, pkg.BorderType.Reflect, pkg.BorderType.Replicate])
    view = big[:, :n0, :]                                      # strides ((n0 + 5) * n1, n1, 1)
    oshape = pkg.conv_device("ndconv_conv_fft", proc, view.data_ptr(), (B, n0, n1), tuple(view.stride()), np.float32, k, pkg.ConvMode.Same, pm, None)
    out = torch.empty(oshape, dtype=torch.float32, device="cuda")
    pkg.conv_device("ndconv_conv_fft", proc, view.data_ptr(), (B, n0, n1), tuple(view.stride()), np.float32, k, pkg.ConvMode.Same, pm, out.data_ptr())
    proc.synchronize()
    pm2 = pkg.PaddingMode.Custom([pkg.BorderType.Reflect, pkg.BorderType.Replicate])
    for b in (0, 3, 6):
        one = pkg.conv_fft_with_processor(view[b].cpu().numpy(), k[0], pkg.ConvMode.Same, pm2, proc)
        assert np.max(np.abs(out[b].cpu().numpy() - one)) <= 2 * fft_tol(np.float32, 1024 * 2048, one, float(np.sum(np.abs(k))))
    proc.close()
