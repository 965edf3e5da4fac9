"""Kernels whose dilated extent exceeds one shared-memory FFT tile on an axis (Kd > 1024 on a strided axis, > 8192 real / 4096
complex on the last; half for f64).  The reference takes any kernel at n log n (fft_size = good_size(max(P, Kd)),
src/conv_fft/mod.rs:229-231); here the kernel is cut into segments along that axis, every segment runs the normal pipeline on the
accordingly shifted / cropped problem and the partial results are summed (conv_fft_split_kernel).  Against the float64 oracle."""
import numpy as np
import pytest

from test_parity_small import fft_tol, mode_from_spec, padding_from_spec

CASES = [
    # shape, kernel, dilation, mode, padding, reverse, dtype
    ((30000,), (9001,), 1, "valid", "zeros", True, np.float32),                                   # 3 segments on the last axis
    ((26000,), (3000,), 4, "full", "reflect", False, np.float32),                                 # dilated: Kd = 11997, segments of 1024 taps
    ((2600, 40), (1500, 3), 1, "same", "replicate", True, np.float32),                            # strided axis: Kd0 = 1500 > 1024
    ((2300, 36), (1100, 2), 1, ("custom", [700, 1], [3, 2]), ("const", 0.75), True, np.float32),  # strides, constant border
    ((9000,), (5000,), 1, "same", ("const", 1.0 - 2.0j), True, np.complex64),                     # complex last axis: cap 4096
    ((1500, 20), (600, 2), 1, "full", "reflect", True, np.float64),                               # f64 strided axis: cap 512
    ((12000,), (9000,), 1, "same", "circular", True, np.float32),                                 # Circular on the cut axis: direct fallback
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}x{c[1]}-{np.dtype(c[6]).name}" for c in CASES])
def test_long_kernel_vs_oracle(ndc, oracle, case):
    pkg, lib = ndc
    shape, ks, dil, mode, padding, rev, dt = case
    rng = np.random.default_rng(21)
    if np.dtype(dt).kind == "c":
        x = (rng.random(shape) - 0.5 + 1j * (rng.random(shape) - 0.5)).astype(dt)
        k = (rng.random(ks) - 0.5 + 1j * (rng.random(ks) - 0.5)).astype(dt)
    else:
        x = (rng.random(shape) - 0.5).astype(dt)
        k = (rng.random(ks) - 0.5).astype(dt)
    kw = pkg.with_dilation(k, dil)
    if not rev:
        kw = kw.no_reverse()
    proc = pkg.get_fft_processor(0, lib)
    l0 = proc.launch_count
    got = pkg.conv_fft_with_processor(x, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), proc)
    launches = proc.launch_count - l0
    ref = oracle.conv_f64_truth(x, k, mode, padding, dil, rev)
    assert got.shape == ref.shape and got.dtype == x.dtype
    tol = fft_tol(dt if np.dtype(dt).kind != "c" else (np.float32 if dt == np.complex64 else np.float64), 8192 * 8, ref,
                  float(np.max(np.abs(x)) * np.sum(np.abs(k))))
    assert np.max(np.abs(got - ref)) <= tol, (float(np.max(np.abs(got - ref))), tol)
    info = pkg.plan_query(shape, dt, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), lib=lib)
    if padding == "circular":
        assert info["path"] == "direct" and launches == 1
    else:
        assert info["path"] == "split" and launches >= 2 * max(info["n_tiles"])      # every segment ran the pipeline


def test_long_kernel_error_order(ndc):
    """validation follows conv_fft's order and codes on this route too: Kd > P -> MismatchShape, an empty kernel -> DataShape"""
    pkg, lib = ndc
    proc = pkg.get_fft_processor(0, lib)
    with pytest.raises(pkg.NdConvError) as e:
        pkg.conv_fft_with_processor(np.ones(5000, np.float32), np.ones(9001, np.float32), pkg.ConvMode.Valid, pkg.PaddingMode.Zeros, proc)
    assert e.value.status == pkg.ERR_MISMATCH_SHAPE
    proc.close()
