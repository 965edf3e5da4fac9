"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/ndconv_oracle.c (the plain-C restatement of
TYPEmber/ndarray-conv's `conv` / `conv_fft`, each C function citing the reference
file:line it follows) plus a scipy/pocketfft restatement of the same pipeline used as the
CPU baseline ("port") by bench.py.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module.  The product package never does.

Parity pinning: see tests/test_oracle_golden.py -- every literal known-answer vector of the
reference's own unit tests, plus torch-CPU generated vectors for its libtorch-derived tests.

Specs (plain Python, mirroring src/lib.rs:80-143):
  mode    : "full" | "same" | "valid" | ("custom", [pad]*N, [stride]*N)
            | ("explicit", [[pf,pb]]*N, [stride]*N)
  padding : "zeros" | ("const", c) | "reflect" | "replicate" | "circular"
            | ("custom", [border]*N) | ("explicit", [[border,border]]*N)
  border  : "zeros" | ("const", c) | "reflect" | "replicate" | "circular"
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"
_LIB = _BUILD / "libndconv_oracle.so"

OK, DATA_SHAPE, KERNEL_SHAPE, MISMATCH_SHAPE, PANIC, BAD_ARG = range(6)
STATUS_NAMES = {0: "Ok", 1: "DataShape", 2: "KernelShape", 3: "MismatchShape", 4: "Panic", 5: "BadArg"}

_DTYPES = {
    np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3,
    np.dtype(np.complex64): 4, np.dtype(np.complex128): 5, np.dtype(np.int8): 6, np.dtype(np.int16): 7,
    np.dtype(np.uint8): 8, np.dtype(np.uint16): 9, np.dtype(np.uint32): 10, np.dtype(np.uint64): 11,
    np.dtype([("lo", "<u8"), ("hi", "<i8")]): 12, np.dtype([("lo", "<u8"), ("hi", "<u8")]): 13,     # Rust i128 / u128 as (lo, hi) records
}
_BORDER = {"zeros": 0, "const": 1, "reflect": 2, "replicate": 3, "circular": 4}
_MODE = {"full": 0, "same": 1, "valid": 2, "custom": 3, "explicit": 4}


class OracleError(Exception):
    def __init__(self, status):
        self.status = status
        super().__init__(STATUS_NAMES.get(status, str(status)))


def build(force: bool = False) -> Path:
    """gcc-compile the C restatement into oracle/_build/ (building the checker is not using it)."""
    srcs = [_HERE / "ndconv_oracle.c", _HERE / "fft_pipeline.inc"]
    if not force and _LIB.exists() and all(_LIB.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return _LIB
    _BUILD.mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(_LIB),
           str(_HERE / "ndconv_oracle.c"), "-lm"]
    subprocess.run(cmd, check=True, cwd=str(_HERE))
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_LIB))
        _lib.orc_good_size_cc.restype = ctypes.c_int64
        _lib.orc_good_size_cc.argtypes = [ctypes.c_int64]
        _lib.orc_gen_offset_list.restype = ctypes.c_int64
    return _lib


def _i64(seq):
    a = np.ascontiguousarray(np.asarray(seq, dtype=np.int64).ravel())
    return a, a.ctypes.data_as(ctypes.c_void_p)


def good_size_cc(n: int) -> int:
    """src/conv_fft/good_size.rs:6-31"""
    return int(lib().orc_good_size_cc(int(n)))


def _mode_args(mode, ndim):
    if isinstance(mode, str):
        return _MODE[mode], np.zeros(2 * ndim, np.int64), np.ones(ndim, np.int64)
    name = mode[0]
    if name == "custom":
        return 3, np.asarray(mode[1], np.int64).ravel(), np.asarray(mode[2], np.int64).ravel()
    if name == "explicit":
        return 4, np.asarray(mode[1], np.int64).ravel(), np.asarray(mode[2], np.int64).ravel()
    raise ValueError(mode)


def unfold(mode, kshape, dilation):
    """ConvMode::unfold, src/conv/mod.rs:28-66 -> (pads[N][2], strides[N])"""
    ndim = len(kshape)
    code, custom, cstr = _mode_args(mode, ndim)
    pads = np.zeros(2 * ndim, np.int64)
    strides = np.zeros(ndim, np.int64)
    ks, ksp = _i64(kshape)
    dl, dlp = _i64(dilation)
    cu, cup = _i64(custom)
    cs, csp = _i64(cstr)
    st = lib().orc_unfold(code, ndim, ksp, dlp, cup, csp, pads.ctypes.data_as(ctypes.c_void_p),
                          strides.ctypes.data_as(ctypes.c_void_p))
    if st:
        raise OracleError(st)
    return pads.reshape(ndim, 2), strides


def lower_padding(padding, ndim, dtype):
    """PaddingMode -> per-side (border codes [N][2], const values [N][2]) the way the
    Custom / Explicit drivers do (src/padding/mod.rs:346-452)."""
    def one(b):
        if isinstance(b, str):
            return _BORDER[b], 0
        return _BORDER[b[0]], b[1]
    codes = np.zeros((ndim, 2), np.int32)
    vals = np.zeros((ndim, 2), dtype)
    if np.dtype(dtype).names:                       # i128 / u128 records: constants are Python ints
        def one(b, _one=one):                       # noqa: F811
            c, v = _one(b)
            u = int(v) % (1 << 128)
            rec = np.zeros((), dtype)
            rec["lo"] = u & ((1 << 64) - 1)
            h = u >> 64
            rec["hi"] = h - (1 << 64) if (np.dtype(dtype)["hi"].kind == "i" and h >= (1 << 63)) else h
            return c, rec
    if isinstance(padding, str) or padding[0] == "const":
        c, v = one(padding)
        codes[:] = c
        vals[:] = v
    elif padding[0] == "custom":
        for i, b in enumerate(padding[1]):
            c, v = one(b)
            codes[i, :] = c
            vals[i, :] = v
    elif padding[0] == "explicit":
        for i, (bf, bb) in enumerate(padding[1]):
            codes[i, 0], vals[i, 0] = one(bf)
            codes[i, 1], vals[i, 1] = one(bb)
    else:
        raise ValueError(padding)
    return codes, vals


def _dilation(d, ndim):
    return [int(d)] * ndim if np.isscalar(d) else [int(v) for v in d]


def pad(x, padding, pads, closed_form=False, buffer_shape=None):
    """PaddingExt::padding (src/padding/mod.rs:84-117); with buffer_shape: the FFT staging
    of conv_fft::padding::data (src/conv_fft/padding.rs:30-62)."""
    x = np.ascontiguousarray(x)
    ndim = x.ndim
    pads = np.asarray(pads, np.int64).reshape(ndim, 2)
    codes, vals = lower_padding(padding, ndim, x.dtype)
    P = [x.shape[i] + int(pads[i, 0] + pads[i, 1]) for i in range(ndim)]
    bshape = list(buffer_shape) if buffer_shape is not None else P
    out = np.zeros(bshape, x.dtype)
    ns, nsp = _i64(x.shape)
    pd, pdp = _i64(pads)
    bs, bsp = _i64(bshape)
    if closed_form:
        assert buffer_shape is None
        st = lib().orc_padding_closed_form(ndim, x.dtype.itemsize, x.ctypes.data_as(ctypes.c_void_p), nsp, pdp,
                                           codes.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p),
                                           out.ctypes.data_as(ctypes.c_void_p))
    else:
        st = lib().orc_padding_in(ndim, x.dtype.itemsize, x.ctypes.data_as(ctypes.c_void_p), nsp, pdp,
                                  codes.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p),
                                  out.ctypes.data_as(ctypes.c_void_p), bsp)
    if st:
        raise OracleError(st)
    return out


def gen_offset_list(kernel, dilation, reverse, pds_strides):
    """KernelWithDilation::gen_offset_list, src/dilation/mod.rs:34-60 -> [(offset, weight)]"""
    k = np.ascontiguousarray(kernel)
    ndim = k.ndim
    offs = np.zeros(max(k.size, 1), np.int64)
    widx = np.zeros(max(k.size, 1), np.int64)
    ks, ksp = _i64(k.shape)
    dl, dlp = _i64(_dilation(dilation, ndim))
    ps, psp = _i64(pds_strides)
    n = lib().orc_gen_offset_list(_DTYPES[k.dtype], ndim, k.ctypes.data_as(ctypes.c_void_p), ksp, dlp, int(bool(reverse)),
                                  psp, offs.ctypes.data_as(ctypes.c_void_p), widx.ctypes.data_as(ctypes.c_void_p))
    flat = k.ravel()
    return [(int(offs[i]), flat[widx[i]]) for i in range(n)]


def _conv_common(fn_name, x, kernel, mode, padding, dilation, reverse, extra):
    x = np.ascontiguousarray(x)
    k = np.ascontiguousarray(kernel, dtype=x.dtype)
    if x.ndim != k.ndim:
        raise ValueError("rank mismatch")
    ndim = x.ndim
    dil = _dilation(dilation, ndim)
    pads, strides = unfold(mode, k.shape, dil)
    codes, vals = lower_padding(padding, ndim, x.dtype)
    ns, nsp = _i64(x.shape)
    ks, ksp = _i64(k.shape)
    dl, dlp = _i64(dil)
    pd, pdp = _i64(pads)
    sd, sdp = _i64(strides)
    oshape = np.zeros(ndim, np.int64)
    fn = getattr(lib(), fn_name)
    args = [_DTYPES[x.dtype], ndim, x.ctypes.data_as(ctypes.c_void_p), nsp, k.ctypes.data_as(ctypes.c_void_p), ksp, dlp,
            int(bool(reverse)), pdp, sdp, codes.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p)]
    st = fn(*args, None, oshape.ctypes.data_as(ctypes.c_void_p), extra)
    if st:
        raise OracleError(st)
    out = np.zeros([int(v) for v in oshape], x.dtype)
    st = fn(*args, out.ctypes.data_as(ctypes.c_void_p), oshape.ctypes.data_as(ctypes.c_void_p), extra)
    if st:
        raise OracleError(st)
    return out


def conv(x, kernel, mode="same", padding="zeros", dilation=1, reverse=True):
    """ConvExt::conv, src/conv/mod.rs:128-200 (ints wrap; floats: mul then add, row-major taps)."""
    return _conv_common("orc_conv_direct", x, kernel, mode, padding, dilation, reverse, 0)


def conv_fft(x, kernel, mode="same", padding="zeros", dilation=1, reverse=True, f64_arith=False):
    """conv_fft_proc_impl, src/conv_fft/mod.rs:185-292, restated in C (native precision, or f64
    arithmetic with f64_arith=True)."""
    return _conv_common("orc_conv_fft", x, kernel, mode, padding, dilation, reverse, 1 if f64_arith else 0)


def conv_f64_truth(x, kernel, mode="same", padding="zeros", dilation=1, reverse=True):
    """Direct convolution evaluated in float64 / complex128: the high-precision yardstick for
    the float tolerance |err| <= c*eps*log2(N)*max|out|."""
    x = np.asarray(x)
    wide = np.complex128 if np.iscomplexobj(x) else np.float64
    if not isinstance(padding, str):
        padding = _widen_padding(padding)
    return conv(x.astype(wide), np.asarray(kernel).astype(wide), mode, padding, dilation, reverse)


def _widen_padding(p):
    return p  # const values are re-cast by lower_padding to the array dtype


# ------------------------------------------------------------------------------------------
# scipy / pocketfft restatement of the reference pipeline: the CPU baseline ("port") that
# bench.py times beside the GPU path.  Same pass structure as src/conv_fft/mod.rs:229-289:
# pad into good_size_cc buffers -> rfftn x2 -> multiply -> irfftn -> crop.  Validated against
# the README numbers within +-15% at survey time.
# ------------------------------------------------------------------------------------------
def conv_fft_scipy(x, kernel, mode="same", padding="zeros", dilation=1, reverse=True, workers=1):
    import scipy.fft as sfft

    x = np.ascontiguousarray(x)
    k = np.ascontiguousarray(kernel, dtype=x.dtype)
    ndim = x.ndim
    dil = _dilation(dilation, ndim)
    pads, strides = unfold(mode, k.shape, dil)
    kd = [k.shape[i] * dil[i] - dil[i] + 1 for i in range(ndim)]
    P = [x.shape[i] + int(pads[i].sum()) for i in range(ndim)]
    F = [good_size_cc(max(P[i], kd[i])) for i in range(ndim)]
    data_pd = pad(x, padding, pads, buffer_shape=F)
    kern_pd = np.zeros(F, x.dtype)
    # conv_fft/padding.rs:98-108: k[j] at j*d (reverse) or at Kd-1-j*d (no_reverse)
    sl = tuple(slice(0, kd[i], dil[i]) if reverse else slice(kd[i] - 1, None, -dil[i]) for i in range(ndim))
    kern_pd[sl] = k
    if np.iscomplexobj(x):
        spec = sfft.fftn(data_pd, workers=workers)
        spec *= sfft.fftn(kern_pd, workers=workers)
        y = sfft.ifftn(spec, workers=workers)
    else:
        spec = sfft.rfftn(data_pd, workers=workers)
        spec *= sfft.rfftn(kern_pd, workers=workers)
        y = sfft.irfftn(spec, s=F, workers=workers)
    crop = tuple(slice(kd[i] - 1, P[i], int(strides[i])) for i in range(ndim))
    return y[crop].astype(x.dtype, copy=False)
