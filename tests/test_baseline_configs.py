"""Parity on the five BASELINE.json configurations at their FULL sizes.

c1-c4: the whole output against the CPU oracle (float64 direct evaluation for the float configs, bit-exact for the i32
direct convolution).  c5 (32768^2, 4.3 GB in / 4.3 GB out) cannot be evaluated by the oracle in reasonable time, so it is
checked through properties the domain offers, on device-resident data:
  * a one-tap kernel turns the convolution into a pure shift of the Reflect-padded input (checked against torch F.pad),
  * random outputs (corners, edges, interior) against float64 direct evaluation of the 63x63 taps,
  * linearity in the data: conv(a*x + b*z) = a*conv(x) + b*conv(z),
  * a constant input gives c * sum(k) everywhere under Reflect;
and, since round 2, three FULL 4096-row bands of the output against the oracle's float64 evaluation of the reference pipeline
(scipy-pocketfft restatement, oracle.conv_fft_scipy) on exactly the input rows those bands read: the top edge, an interior band
across tile seams, and the bottom band that holds the axis-0 tile seams 29822 / 30784 / 31746, the axis-0 split row 32708 and
the bottom Reflect edge.
"""
import numpy as np
import pytest

from test_parity_small import fft_tol

pytestmark = pytest.mark.gpu


def synth(c, shape, dtype=np.float32, kernel=False):
    rng = np.random.default_rng((2000 if kernel else 1000) + c)
    if dtype == np.int32:
        return rng.integers(-128, 128, size=shape, dtype=np.int32)
    if dtype == np.complex64:
        return (rng.random(shape, dtype=np.float32) + 1j * rng.random(shape, dtype=np.float32)).astype(np.complex64)
    return rng.random(shape, dtype=np.float32)


def test_c1_1d_f32(pkg, cuda_lib, oracle):
    x, k = synth(1, (5000,)), synth(1, (31,), kernel=True)
    got = pkg.conv_fft(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Zeros, lib=cuda_lib)
    ref = oracle.conv_f64_truth(x, k, "same", "zeros")
    assert got.shape == (5000,) and np.max(np.abs(got - ref)) <= fft_tol(np.float32, 512, ref)


def test_c2_2d_f32_dilated_reflect_circular(pkg, cuda_lib, oracle):
    x, k = synth(2, (200, 5000)), synth(2, (11, 31), kernel=True)
    pm = pkg.PaddingMode.Custom([pkg.BorderType.Reflect, pkg.BorderType.Circular])
    got = pkg.conv_fft(x, pkg.with_dilation(k, 2), pkg.ConvMode.Same, pm, lib=cuda_lib)
    ref = oracle.conv_f64_truth(x, k, "same", ("custom", ["reflect", "circular"]), 2)
    assert got.shape == (200, 5000) and np.max(np.abs(got - ref)) <= fft_tol(np.float32, 256 * 2048, ref)


@pytest.mark.parametrize("dtype", [np.float32, np.complex64], ids=["f32", "c32"])
def test_c3_3d(pkg, cuda_lib, oracle, dtype):
    x, k = synth(3, (10, 100, 200), dtype), synth(3, (5, 11, 31), dtype, kernel=True)
    got = pkg.conv_fft(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Zeros, lib=cuda_lib)
    ref = oracle.conv_f64_truth(x, k, "same", "zeros")
    assert got.shape == (10, 100, 200) and np.max(np.abs(got - ref)) <= fft_tol(dtype, 16 * 128 * 256, ref)
    par = pkg.conv_fft_par(x, k, pkg.ConvMode.Same, pkg.PaddingMode.Zeros, lib=cuda_lib)
    assert np.array_equal(par, got)


def test_c4_3d_direct_i32_bit_exact(pkg, cuda_lib, oracle):
    x, k = synth(4, (64, 256, 256), np.int32), synth(4, (3, 5, 5), np.int32, kernel=True)
    got = pkg.conv(x, k, pkg.ConvMode.Custom([1, 2, 2], [2, 2, 2]), pkg.PaddingMode.Replicate, lib=cuda_lib)
    ref = oracle.conv(x, k, ("custom", [1, 2, 2], [2, 2, 2]), "replicate")
    assert got.shape == (32, 128, 128) and np.array_equal(got, ref)
    # the same problem in f32 / i64: still bit-identical (same tap order, un-fused multiply-add)
    for dt in (np.float32, np.int64):
        g2 = pkg.conv(x.astype(dt), k.astype(dt), pkg.ConvMode.Custom([1, 2, 2], [2, 2, 2]), pkg.PaddingMode.Replicate, lib=cuda_lib)
        r2 = oracle.conv(x.astype(dt), k.astype(dt), ("custom", [1, 2, 2], [2, 2, 2]), "replicate")
        assert g2.tobytes() == r2.tobytes()


def test_c5_full_size_properties(pkg, cuda_lib):
    torch = pytest.importorskip("torch")
    import torch.nn.functional as F
    n, K = 32768, 63
    dev = torch.device("cuda", 0)
    proc = pkg.get_fft_processor(0, cuda_lib)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    proc.set_stream(stream.cuda_stream)
    g = torch.Generator(device=dev).manual_seed(1005)
    x = torch.rand((n, n), generator=g, device=dev, dtype=torch.float32)
    y = torch.empty((n + K - 1, n + K - 1), device=dev, dtype=torch.float32)

    def run(xin, k, out):
        pkg.conv_device("ndconv_conv_fft", proc, xin.data_ptr(), (n, n), (n, 1), np.float32, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, out.data_ptr())
        torch.cuda.synchronize(dev)

    eps = float(np.finfo(np.float32).eps)
    tol_rel = 4 * eps * np.log2(1024 * 2048)
    # (1) one-tap kernel: out[o] = B[o + (K-1) - (a,b)], B = Reflect-padded x
    a, b = 17, 40
    k1 = np.zeros((K, K), np.float32)
    k1[a, b] = 1.0
    run(x, k1, y)
    pad = K - 1
    for r0 in (0, 9000, n + K - 1 - 4096):                       # check three 4096-row bands (top edge, interior, bottom edge)
        rows = torch.arange(r0, r0 + 4096, device=dev) + (K - 1 - a) - pad     # source row before reflection
        rows = torch.where(rows < 0, -rows, rows)
        rows = torch.where(rows >= n, 2 * (n - 1) - rows, rows)
        band = F.pad(x.index_select(0, rows)[None, None], [pad, pad, 0, 0], mode="reflect")[0, 0]     # reflect along axis 1
        expect = band[:, (K - 1 - b):(K - 1 - b) + n + K - 1]
        assert float((y[r0:r0 + 4096] - expect).abs().max()) <= tol_rel * 1.0
    # (2) random taps: spot checks against float64 direct evaluation
    kr = np.random.default_rng(2005).random((K, K), dtype=np.float32)
    run(x, kr, y)
    rng = np.random.default_rng(3)
    O = n + K - 1
    pts = [(0, 0), (O - 1, O - 1), (0, O - 1), (O - 1, 0), (61, 61), (62, 62)] + [(int(rng.integers(0, O)), int(rng.integers(0, O))) for _ in range(120)]
    kf = torch.from_numpy(kr[::-1, ::-1].astype(np.float64).copy()).to(dev)
    worst, scale = 0.0, 0.0
    ar = torch.arange(K, device=dev)
    for (o0, o1) in pts:
        r = ar + o0 - pad
        c = ar + o1 - pad
        r = torch.where(r < 0, -r, r); r = torch.where(r >= n, 2 * (n - 1) - r, r)
        c = torch.where(c < 0, -c, c); c = torch.where(c >= n, 2 * (n - 1) - c, c)
        ref = float((x.index_select(0, r).index_select(1, c).double() * kf).sum())
        worst = max(worst, abs(float(y[o0, o1]) - ref)); scale = max(scale, abs(ref))
    assert worst <= tol_rel * scale, (worst, scale)
    # (3) constant input: every output equals c * sum(k)
    xc = torch.full((n, n), 0.75, device=dev, dtype=torch.float32)
    y2 = torch.empty_like(y)
    run(xc, kr, y2)
    target = 0.75 * float(kr.astype(np.float64).sum())
    assert float((y2 - target).abs().max()) <= tol_rel * abs(target)
    # (4) linearity in the data (y holds conv(x)): conv(0.5 x + 2 xc) = 0.5 conv(x) + 2 conv(xc)
    xc.mul_(2.0).add_(x, alpha=0.5)
    y3 = torch.empty((4096, O), device=dev, dtype=torch.float32)
    full = torch.empty_like(y2)
    lin = y[:4096] * 0.5 + y2[:4096] * 2.0
    run(xc, kr, full)
    assert float((full[:4096] - lin).abs().max()) <= 3 * tol_rel * float(lin.abs().max())
    del y3
    proc.close()


def test_c5_full_bands_against_oracle(pkg, cuda_lib, oracle):
    """c5 at full size on the device; bands of 4096 output rows (all 32830 columns) against the float64 oracle.  A band of output
    rows [r0, r0 + B) reads padded rows [r0, r0 + B + Kd0 - 1) of axis 0; the Reflect border of axis 0 is resolved here by index
    (rows of the device array are fetched through the reflected index), and the oracle then runs the band as
    ConvMode::Explicit{axis 0: no padding, axis 1: 62 | 62} with Reflect on axis 1 -- the overlap-save identity of SURVEY 8e."""
    torch = pytest.importorskip("torch")
    n, K = 32768, 63
    dev = torch.device("cuda", 0)
    proc = pkg.get_fft_processor(0, cuda_lib)
    g = torch.Generator(device=dev).manual_seed(1005)
    x = torch.rand((n, n), generator=g, device=dev, dtype=torch.float32)
    O = n + K - 1
    y = torch.empty((O, O), device=dev, dtype=torch.float32)
    kr = np.random.default_rng(2005).random((K, K), dtype=np.float32)
    torch.cuda.synchronize(dev)
    pkg.conv_device("ndconv_conv_fft", proc, x.data_ptr(), (n, n), (n, 1), np.float32, kr, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, y.data_ptr())
    proc.synchronize()
    info = pkg.plan_query((n, n), np.float32, kr, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, lib=cuda_lib)
    V0 = info["tile_valid"][0]
    B = 4096
    seam = (17 * V0) - B // 2                                  # an interior band with axis-0 tile seams inside it
    eps = float(np.finfo(np.float32).eps)
    for r0 in (0, seam, O - B):
        pr = torch.arange(r0, r0 + B + K - 1, device=dev) - (K - 1)            # source row of each padded row, before reflection
        pr = torch.where(pr < 0, -pr, pr)
        pr = torch.where(pr >= n, 2 * (n - 1) - pr, pr)
        xb = x.index_select(0, pr).cpu().numpy().astype(np.float64)
        ref = oracle.conv_fft_scipy(xb, kr.astype(np.float64), ("explicit", [[0, 0], [K - 1, K - 1]], [1, 1]),
                                    ("explicit", [["zeros", "zeros"], ["reflect", "reflect"]]), 1, True, workers=__import__("os").cpu_count() or 1)
        got = y[r0:r0 + B].cpu().numpy().astype(np.float64)
        assert ref.shape == got.shape == (B, O)
        scale = float(np.max(np.abs(ref)))
        err = float(np.max(np.abs(got - ref)))
        tol = 4 * eps * np.log2(1024 * 2048) * scale
        print(f"c5 band rows [{r0}, {r0 + B}): max|err| = {err:.3e}, tol = {tol:.3e}, max|out| = {scale:.1f}")
        assert err <= tol, (r0, err, tol)
    proc.close()
