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
# outside the shared-memory envelope -> global-memory Stockham passes: odd real axes, prime factors above 7 (rustfft takes any
# length: real.rs:40,62, complex.rs:56), axes longer than one shared-memory transform
ANY_LENGTH_SHAPES = [(1,), (7,), (22,), (13,), (97,), (1, 1), (11, 13), (6, 15), (26, 34), (3, 5, 7), (2, 11, 9), (17, 4, 6), (3, 2, 5, 3),
                     (1100, 6), (3, 1056, 4), (2, 9000), (8400,),
                     # prime factors above 97: chirp-z (Bluestein) through power-of-two passes instead of an O(n r) generic pass
                     (101,), (2, 1009), (211, 4), (3, 202, 5), (2018,)]


@pytest.mark.parametrize("shape", SHAPES, ids=str)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128], ids=["f32", "f64", "c32", "c64"])
def test_forward_layout_and_roundtrip(ndc, shape, dtype):
    pkg, lib = ndc
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape).astype(dtype)
    if np.iscomplexobj(x):
        x = x + 1j * rng.standard_normal(shape).astype(dtype)
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


@pytest.mark.parametrize("shape", ANY_LENGTH_SHAPES, ids=str)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128], ids=["f32", "f64", "c32", "c64"])
def test_any_length(ndc, shape, dtype):
    test_forward_layout_and_roundtrip(ndc, shape, dtype)


def test_spectrum_of_other_origin_len(ndc):
    """real.rs:41,108: backward takes the real-space last-axis length from the processor (rp_origin_len), so the same half
    spectrum width (6 bins) inverts to 10 or to 11 samples depending on the shape handed over."""
    pkg, lib = ndc
    proc = pkg.get_fft_processor(0, lib)
    rng = np.random.default_rng(3)
    for n in (10, 11):
        x = rng.standard_normal((4, n))
        spec = np.moveaxis(np.fft.rfftn(x), 0, -1)
        assert spec.shape == (6, 4)
        back = proc.backward(spec, shape=(4, n), dtype=np.float64)
        assert np.max(np.abs(back - x)) <= 1e-10
    proc.close()


def test_random_shapes_sweep(ndc):
    """seeded random ranks 1-4, lengths mixing small, prime, smooth and longer-than-one-tile axes, all four element types:
    forward against numpy.fft (rotated layout) and forward o backward = id"""
    pkg, lib = ndc
    rng = np.random.default_rng(11)
    proc = pkg.get_fft_processor(0, lib)
    done = 0
    while done < 120:
        nd = int(rng.integers(1, 5))
        lens = []
        for _ in range(nd):
            c = int(rng.integers(0, 4))
            if c == 0:
                lens.append(int(rng.integers(1, 40)))
            elif c == 1:
                lens.append(int(rng.choice([11, 13, 17, 19, 23, 29, 31, 37, 41, 53, 64, 81, 100, 121, 127])))
            elif c == 2:
                lens.append(int(rng.integers(1, 12)))
            else:
                lens.append(int(rng.choice([1, 2, 3, 1030, 1100, 2058][: 3 if nd > 2 else 6])))
        if np.prod(lens) > 400000:
            continue
        done += 1
        dt = [np.float32, np.float64, np.complex64, np.complex128][int(rng.integers(0, 4))]
        x = rng.standard_normal(lens).astype(dt)
        if np.iscomplexobj(x):
            x = x + 1j * rng.standard_normal(lens).astype(dt)
        spec = proc.forward(x)
        ref = expected_spectrum(x)
        eps = np.finfo(np.float32 if dt in (np.float32, np.complex64) else np.float64).eps
        tol = 8 * eps * np.log2(max(x.size, 2))
        assert spec.shape == ref.shape, lens
        assert np.max(np.abs(spec - ref)) <= tol * max(np.max(np.abs(ref)), 1e-30), (lens, dt)
        back = proc.backward(spec)
        assert np.max(np.abs(back - x)) <= tol * max(np.max(np.abs(x)), 1e-30), (lens, dt)
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


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((2048, 2048), np.float32), ((65536,), np.float32), ((4099,), np.complex64), ((3000, 1001), np.float64),
                                          ((6, 1500, 130), np.complex64)], ids=str)
def test_global_passes_at_size(pkg, cuda_lib, shape, dtype):
    """axes outside the shared-memory envelope at sizes that fill the device: 2048-point strided axis, a 65536-point row,
    a prime length, an odd real axis next to a 3000-point strided axis, a 1500-point middle axis"""
    rng = np.random.default_rng(4)
    x = rng.standard_normal(shape).astype(dtype)
    if np.iscomplexobj(x):
        x = x + 1j * rng.standard_normal(shape).astype(dtype)
    proc = pkg.get_fft_processor(0, cuda_lib)
    spec = proc.forward(x)
    ref = expected_spectrum(x)
    assert spec.shape == ref.shape
    eps = np.finfo(np.float32 if x.dtype in (np.float32, np.complex64) else np.float64).eps
    assert np.max(np.abs(spec - ref)) <= 8 * eps * np.log2(x.size) * np.max(np.abs(ref))
    back = proc.backward(spec)
    assert np.max(np.abs(back - x)) <= (1e-6 if eps > 1e-10 else 1e-10) * max(1.0, np.max(np.abs(x))) * 4
    proc.close()
