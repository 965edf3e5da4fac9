"""Processor::forward / backward (src/conv_fft/processor/{real,complex}.rs): unnormalised N-d FFT with the reference's rotated
spectrum layout (axis 0 moves to the end), inverse scaled by 1/len.  Checked against numpy.fft (float64) and
through the reference's own round-trip property (real.rs:291-478, complex.rs:154-263: forward o backward = id, 1e-6 f32 /
1e-10 f64)."""
import numpy as np
import pytest


def expected_spectrum(x):
    if np.iscomplexobj(x):
        s = np.fft.fftn(x.astype(np.complex128))
    else:
        s = np.fft.rfftn(x.astype(np.float64))
    return np.moveaxis(s, 0, -1) if x.ndim > 1 else s


SHAPES = [(16,), (30,), (6, 10), (12, 40), (4, 6, 8), (5, 7, 12), (3, 4, 5, 6)]


@pytest.mark.parametrize("shape", SHAPES, ids=str)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128], ids=["f32", "f64", "c32", "c64"])
def test_forward_layout_and_roundtrip(ndc, shape, dtype):
    pkg, lib = ndc
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape).astype(dtype)
    if np.iscomplexobj(x):
        x = x + 1j * rng.standard_normal(shape).astype(dtype)
    elif shape[-1] % 2:
        pytest.skip("odd real last axis is outside this build's envelope")
    proc = pkg.get_fft_processor(0, lib)
    spec = proc.forward(x)
    ref = expected_spectrum(x)
    assert spec.shape == ref.shape
    eps = np.finfo(np.float32 if x.dtype.itemsize in (4, 8) and x.dtype in (np.float32, np.complex64) else np.float64).eps
    assert np.max(np.abs(spec - ref)) <= 8 * eps * np.log2(max(x.size, 2)) * np.max(np.abs(ref))
    back = proc.backward(spec)
    assert back.shape == x.shape and back.dtype == x.dtype
    assert np.max(np.abs(back - x)) <= (1e-6 if eps > 1e-10 else 1e-10) * max(1.0, np.max(np.abs(x)))
    proc.close()


def test_unsupported_lengths_are_reported(ndc):
    pkg, lib = ndc
    proc = pkg.get_fft_processor(0, lib)
    with pytest.raises(pkg.NdConvError) as e:
        proc.forward(np.zeros(22, np.float32))      # 11 is not {2,3,5,7}-smooth
    assert e.value.status == pkg.ERR_UNSUPPORTED
    with pytest.raises(pkg.NdConvError) as e:
        proc.forward(np.zeros((2000, 4), np.float32))   # strided axis longer than one shared-memory transform
    assert e.value.status == pkg.ERR_UNSUPPORTED
    proc.close()


@pytest.mark.gpu
def test_reference_roundtrip_200x5000(pkg, cuda_lib):
    """real.rs:405-447: 200 x 5000 f32 round trip, tolerance 1e-6"""
    rng = np.random.default_rng(2)
    x = rng.random((200, 5000), dtype=np.float32)
    proc = pkg.get_fft_processor(0, cuda_lib)
    spec = proc.forward(x)
    assert spec.shape == (2501, 200)
    ref = expected_spectrum(x)
    assert np.max(np.abs(spec - ref)) <= 8 * np.finfo(np.float32).eps * np.log2(x.size) * np.max(np.abs(ref))
    back = proc.backward(spec)
    assert np.max(np.abs(back - x)) < 1e-6
    proc.close()
