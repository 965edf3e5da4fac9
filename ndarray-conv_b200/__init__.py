"""ndarray-conv_b200 -- host-side mirror (Python) of TYPEmber/ndarray-conv's trait API over the
C ABI of libndconv_cuda.so (include/ndconv.h).

The reference is a Rust crate; this image has no Rust toolchain, so the drop-in boundary is the
C ABI and the host mirrors are C++ (include/ndconv.hpp here) and this ctypes module, which the
parity tests and bench.py drive.  Names follow the reference:

    ConvMode.Full / Same / Valid / Custom(padding, strides) / Explicit(padding, strides)   src/lib.rs:80-105
    PaddingMode.Zeros / Const(c) / Reflect / Replicate / Circular / Custom([...]) / Explicit([[..,..]])  src/lib.rs:111-127
    BorderType.Zeros / Const(c) / Reflect / Replicate / Circular                            src/lib.rs:131-143
    with_dilation(kernel, d) [.no_reverse() / .reverse()]                                    src/dilation/mod.rs:103-189
    conv(x, kernel, conv_mode, padding_mode)                                                 src/conv/mod.rs:110-115
    conv_fft / conv_fft_with_processor / conv_fft_par                                        src/conv_fft/mod.rs:113-171
    get_fft_processor()                                                                      src/conv_fft/processor/mod.rs:71-73
    NdConvError(DataShape | KernelShape | MismatchShape)                                     src/lib.rs:148-159

There is no CPU fallback: importing works anywhere, but any compute call raises unless the CUDA
library is built (python __graft_entry__.py / ndarray-conv_b200/build.py) and a B200 is present.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libndconv_cuda.so"
MAX_DIM = 6

OK, ERR_DATA_SHAPE, ERR_KERNEL_SHAPE, ERR_MISMATCH_SHAPE, ERR_PANIC, ERR_BAD_ARG, ERR_UNSUPPORTED = range(7)
ERR_CUDA, ERR_INTERNAL = 100, 101
PATH_DIRECT, PATH_FFT = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1

# Rust i128 / u128 have no numpy scalar type: 16-byte little-endian records (lo, hi); int128_array / int128_values convert
I128 = np.dtype([("lo", "<u8"), ("hi", "<i8")])
U128 = np.dtype([("lo", "<u8"), ("hi", "<u8")])


def int128_array(values, signed=True):
    """nested sequence (or numpy array) of Python ints -> array of I128 / U128 records (two's complement, wrapping mod 2^128)"""
    obj = np.asarray(values, dtype=object)
    out = np.zeros(obj.shape, I128 if signed else U128)
    flat_lo, flat_hi = out["lo"].reshape(-1), out["hi"].reshape(-1)
    for i, v in enumerate(obj.reshape(-1)):
        u = int(v) % (1 << 128)
        flat_lo[i] = u & ((1 << 64) - 1)
        h = u >> 64
        flat_hi[i] = h - (1 << 64) if (signed and h >= (1 << 63)) else h
    return out


def int128_values(arr):
    """array of I128 / U128 records -> object array of Python ints"""
    arr = np.asarray(arr)
    signed = arr.dtype == I128
    lo, hi = arr["lo"].astype(object), arr["hi"].astype(object)
    v = (hi % (1 << 64)) * (1 << 64) + lo
    if signed:
        v = np.where(v >= (1 << 127), v - (1 << 128), v)
    return v


def _scalar_bytes(value, dtype):
    if dtype in (I128, U128):
        return (int(value) % (1 << 128)).to_bytes(16, "little")
    return np.asarray(value, dtype=dtype).tobytes()


DTYPE_CODES = {
    np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3,
    np.dtype(np.complex64): 4, np.dtype(np.complex128): 5, np.dtype(np.int8): 6, np.dtype(np.int16): 7,
    np.dtype(np.uint8): 8, np.dtype(np.uint16): 9, np.dtype(np.uint32): 10, np.dtype(np.uint64): 11,
    I128: 12, U128: 13,
}


class NdConvError(Exception):
    """Mirror of `Error<N>` (src/lib.rs:148-159) plus the non-reference statuses of include/ndconv.h."""

    NAMES = {1: "DataShape", 2: "KernelShape", 3: "MismatchShape", 4: "ReferencePanic", 5: "BadArgument",
             6: "Unsupported", 100: "CudaError", 101: "InternalError"}

    def __init__(self, status, message=""):
        self.status = status
        self.kind = self.NAMES.get(status, str(status))
        super().__init__(f"{self.kind}: {message}")


# ----------------------------------------------------------------------------------------------
# ctypes mirror of include/ndconv.h
# ----------------------------------------------------------------------------------------------
class _Border(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("reserved", ctypes.c_int32), ("value", ctypes.c_ubyte * 16)]


class _Problem(ctypes.Structure):
    _fields_ = [
        ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32), ("memory", ctypes.c_int32), ("reverse", ctypes.c_int32),
        ("data", ctypes.c_void_p), ("data_shape", ctypes.c_int64 * MAX_DIM), ("data_strides", ctypes.c_int64 * MAX_DIM),
        ("kernel", ctypes.c_void_p), ("kernel_shape", ctypes.c_int64 * MAX_DIM), ("kernel_strides", ctypes.c_int64 * MAX_DIM),
        ("dilation", ctypes.c_int64 * MAX_DIM),
        ("pad", (ctypes.c_int64 * 2) * MAX_DIM),
        ("stride", ctypes.c_int64 * MAX_DIM),
        ("border", (_Border * 2) * MAX_DIM),
    ]


class _PlanInfo(ctypes.Structure):
    _fields_ = [("path", ctypes.c_int), ("ndim", ctypes.c_int), ("tile_len", ctypes.c_int * 6), ("tile_valid", ctypes.c_int * 6),
                ("n_tiles", ctypes.c_int * 6), ("workspace_bytes", ctypes.c_int64), ("split_out_rows", ctypes.c_int64), ("pipelined", ctypes.c_int)]


class _Slab(ctypes.Structure):
    _fields_ = [("out_begin", ctypes.c_int64), ("out_end", ctypes.c_int64), ("pad_begin", ctypes.c_int64), ("pad_end", ctypes.c_int64)]


class _Shard(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("rows", ctypes.c_int64), ("halo_front", ctypes.c_int64), ("halo_back", ctypes.c_int64), ("out", ctypes.c_void_p)]


class _ShardInfo(ctypes.Structure):
    _fields_ = [("out_begin", ctypes.c_int64), ("out_end", ctypes.c_int64), ("halo_front", ctypes.c_int64), ("halo_back", ctypes.c_int64),
                ("first_row", ctypes.c_int64)]


# every symbol include/ndconv.h declares (tests check the built library exports all of them)
EXPORTED_SYMBOLS = [
    "ndconv_version", "ndconv_is_emulation", "ndconv_last_error_string", "ndconv_status_string", "ndconv_dtype_size",
    "ndconv_device_count", "ndconv_unfold_conv_mode", "ndconv_good_fft_size", "ndconv_plan_fft_size", "ndconv_out_shape",
    "ndconv_border_index_map", "ndconv_processor_create", "ndconv_processor_destroy", "ndconv_processor_set_stream",
    "ndconv_processor_synchronize", "ndconv_processor_launch_count", "ndconv_processor_workspace_bytes",
    "ndconv_processor_set_profiling", "ndconv_processor_get_profile",
    "ndconv_conv_direct", "ndconv_conv_fft", "ndconv_conv_fft_par", "ndconv_conv_fft_sharded", "ndconv_shard_plan", "ndconv_conv_fft_sharded_device", "ndconv_conv_fft_batch", "ndconv_fft_forward", "ndconv_fft_backward",
    "ndconv_plan_query", "ndconv_slab_plan", "ndconv_host_alloc", "ndconv_host_free", "ndconv_host_register", "ndconv_host_unregister",
]


class Library:
    """Typed binding of one loaded libndconv shared object."""

    def __init__(self, cdll):
        self.c = cdll
        c = cdll
        c.ndconv_version.restype = ctypes.c_char_p
        c.ndconv_last_error_string.restype = ctypes.c_char_p
        c.ndconv_status_string.restype = ctypes.c_char_p
        c.ndconv_status_string.argtypes = [ctypes.c_int]
        c.ndconv_dtype_size.restype = ctypes.c_size_t
        c.ndconv_dtype_size.argtypes = [ctypes.c_int]
        c.ndconv_good_fft_size.restype = ctypes.c_int64
        c.ndconv_good_fft_size.argtypes = [ctypes.c_int64]
        c.ndconv_plan_fft_size.restype = ctypes.c_int64
        c.ndconv_plan_fft_size.argtypes = [ctypes.c_int64, ctypes.c_int]
        c.ndconv_out_shape.argtypes = [ctypes.POINTER(_Problem), ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
        c.ndconv_border_index_map.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        c.ndconv_unfold_conv_mode.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        c.ndconv_processor_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        c.ndconv_processor_destroy.argtypes = [ctypes.c_void_p]
        c.ndconv_processor_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        c.ndconv_processor_synchronize.argtypes = [ctypes.c_void_p]
        c.ndconv_processor_launch_count.restype = ctypes.c_int64
        c.ndconv_processor_launch_count.argtypes = [ctypes.c_void_p]
        c.ndconv_processor_workspace_bytes.restype = ctypes.c_int64
        c.ndconv_processor_workspace_bytes.argtypes = [ctypes.c_void_p]
        c.ndconv_processor_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
        c.ndconv_processor_get_profile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        for name in ("ndconv_conv_direct", "ndconv_conv_fft", "ndconv_conv_fft_par"):
            getattr(c, name).argtypes = [ctypes.c_void_p, ctypes.POINTER(_Problem), ctypes.c_void_p]
        c.ndconv_conv_fft_sharded.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(_Problem), ctypes.c_void_p]
        c.ndconv_conv_fft_batch.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(_Problem), ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        for name in ("ndconv_fft_forward", "ndconv_fft_backward"):
            getattr(c, name).argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        c.ndconv_plan_query.argtypes = [ctypes.POINTER(_Problem), ctypes.POINTER(_PlanInfo)]
        c.ndconv_shard_plan.argtypes = [ctypes.POINTER(_Problem), ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.POINTER(_ShardInfo)]
        c.ndconv_conv_fft_sharded_device.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(_Problem), ctypes.POINTER(_Shard)]
        c.ndconv_slab_plan.argtypes = [ctypes.POINTER(_Problem), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_Slab)]
        c.ndconv_host_alloc.restype = ctypes.c_void_p
        c.ndconv_host_alloc.argtypes = [ctypes.c_size_t]
        c.ndconv_host_free.argtypes = [ctypes.c_void_p]
        c.ndconv_host_register.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        c.ndconv_host_unregister.argtypes = [ctypes.c_void_p]

    def check(self, status):
        if status != 0:
            raise NdConvError(status, (self.c.ndconv_last_error_string() or b"").decode())

    @property
    def is_emulation(self):
        return bool(self.c.ndconv_is_emulation())


_library = None


def get_library() -> Library:
    """Load the CUDA product library.  Fails loudly when it has not been built; never substitutes anything else."""
    global _library
    if _library is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                               "(nvcc, sm_100a). There is no CPU fallback.")
        lib = Library(ctypes.CDLL(str(LIB_PATH)))
        if lib.is_emulation:
            raise RuntimeError("refusing to load a host-emulation build as the product library")
        _library = lib
    return _library


# ----------------------------------------------------------------------------------------------
# the reference's enums
# ----------------------------------------------------------------------------------------------
class BorderType:
    """src/lib.rs:131-143"""
    ZEROS, CONST, REFLECT, REPLICATE, CIRCULAR = range(5)

    def __init__(self, kind, value=0):
        self.kind, self.value = kind, value

    @classmethod
    def Const(cls, c):
        return cls(cls.CONST, c)

    def spec(self):
        name = ["zeros", "const", "reflect", "replicate", "circular"][self.kind]
        return ("const", self.value) if self.kind == self.CONST else name


BorderType.Zeros = BorderType(BorderType.ZEROS)
BorderType.Reflect = BorderType(BorderType.REFLECT)
BorderType.Replicate = BorderType(BorderType.REPLICATE)
BorderType.Circular = BorderType(BorderType.CIRCULAR)


class PaddingMode:
    """src/lib.rs:111-127.  `lower(ndim)` performs the per-side lowering of the Custom / Explicit
    drivers (src/padding/mod.rs:346-452); the five plain modes use the same border on every side."""

    def __init__(self, kind, payload=None):
        self.kind, self.payload = kind, payload

    @classmethod
    def Const(cls, c):
        return cls("const", c)

    @classmethod
    def Custom(cls, borders):
        return cls("custom", list(borders))

    @classmethod
    def Explicit(cls, borders):
        return cls("explicit", [list(b) for b in borders])

    def lower(self, ndim):
        plain = {"zeros": BorderType.Zeros, "reflect": BorderType.Reflect, "replicate": BorderType.Replicate, "circular": BorderType.Circular}
        if self.kind in plain:
            return [[plain[self.kind]] * 2 for _ in range(ndim)]
        if self.kind == "const":
            return [[BorderType.Const(self.payload)] * 2 for _ in range(ndim)]
        if self.kind == "custom":
            assert len(self.payload) == ndim
            return [[b, b] for b in self.payload]
        assert len(self.payload) == ndim
        return [[b[0], b[1]] for b in self.payload]

    def spec(self):
        """the plain-tuple form oracle/oracle.py takes (tests only)"""
        if self.kind == "const":
            return ("const", self.payload)
        if self.kind == "custom":
            return ("custom", [b.spec() for b in self.payload])
        if self.kind == "explicit":
            return ("explicit", [[b[0].spec(), b[1].spec()] for b in self.payload])
        return self.kind


PaddingMode.Zeros = PaddingMode("zeros")
PaddingMode.Reflect = PaddingMode("reflect")
PaddingMode.Replicate = PaddingMode("replicate")
PaddingMode.Circular = PaddingMode("circular")


class ConvMode:
    """src/lib.rs:80-105"""
    FULL, SAME, VALID, CUSTOM, EXPLICIT = range(5)

    def __init__(self, kind, padding=None, strides=None):
        self.kind, self.padding, self.strides = kind, padding, strides

    @classmethod
    def Custom(cls, padding, strides):
        return cls(cls.CUSTOM, list(padding), list(strides))

    @classmethod
    def Explicit(cls, padding, strides):
        return cls(cls.EXPLICIT, [list(p) for p in padding], list(strides))

    def spec(self):
        if self.kind == self.CUSTOM:
            return ("custom", self.padding, self.strides)
        if self.kind == self.EXPLICIT:
            return ("explicit", self.padding, self.strides)
        return ["full", "same", "valid"][self.kind]

    def unfold(self, kernel_shape, dilation, lib: Library | None = None):
        """ConvMode::unfold (src/conv/mod.rs:28-66) through the C ABI -> (pads [N][2], strides [N])"""
        lib = lib or get_library()
        n = len(kernel_shape)
        ks = (ctypes.c_int64 * n)(*kernel_shape)
        dl = (ctypes.c_int64 * n)(*dilation)
        pad = np.zeros((n, 2), np.int64)
        strd = np.zeros(n, np.int64)
        if self.kind == self.CUSTOM:
            pp = np.asarray(self.padding, np.int64).reshape(n)
        elif self.kind == self.EXPLICIT:
            pp = np.asarray(self.padding, np.int64).reshape(n, 2)
        else:
            pp = np.zeros(n, np.int64)
        ss = np.asarray(self.strides if self.strides is not None else [1] * n, np.int64)
        pp = np.ascontiguousarray(pp)
        lib.check(lib.c.ndconv_unfold_conv_mode(self.kind, n, ks, dl, pp.ctypes.data, ss.ctypes.data, pad.ctypes.data, strd.ctypes.data))
        return pad, strd


ConvMode.Full = ConvMode(ConvMode.FULL)
ConvMode.Same = ConvMode(ConvMode.SAME)
ConvMode.Valid = ConvMode(ConvMode.VALID)


class KernelWithDilation:
    """src/dilation/mod.rs:8-12: kernel + dilation[N] + reverse (default True = mathematical convolution)."""

    def __init__(self, kernel, dilation, reverse=True):
        self.kernel = np.asarray(kernel)
        nd = self.kernel.ndim
        self.dilation = [int(dilation)] * nd if np.isscalar(dilation) else [int(d) for d in dilation]
        self.reverse_flag = reverse

    def reverse(self):
        return KernelWithDilation(self.kernel, self.dilation, True)

    def no_reverse(self):
        return KernelWithDilation(self.kernel, self.dilation, False)


def with_dilation(kernel, dilation) -> KernelWithDilation:
    """WithDilation::with_dilation, src/dilation/mod.rs:103-126"""
    return KernelWithDilation(kernel, dilation, True)


def _into_kwd(kernel) -> KernelWithDilation:
    """IntoKernelWithDilation, src/dilation/mod.rs:192-212: a bare array means dilation 1, reverse"""
    return kernel if isinstance(kernel, KernelWithDilation) else KernelWithDilation(kernel, 1, True)


# ----------------------------------------------------------------------------------------------
# problem lowering
# ----------------------------------------------------------------------------------------------
def make_problem(x_shape, x_strides_elems, x_ptr, dtype, kernel: KernelWithDilation, conv_mode: ConvMode,
                 padding_mode: PaddingMode, memory=MEM_HOST, lib: Library | None = None, explicit=None):
    """Build an ndconv_problem.  Returns (problem, keepalive).  `explicit` = (pads, strides) skips unfold."""
    lib = lib or get_library()
    dtype = np.dtype(dtype)
    if dtype not in DTYPE_CODES:
        raise NdConvError(ERR_BAD_ARG, f"unsupported dtype {dtype}")
    nd = len(x_shape)
    k = np.asarray(kernel.kernel)
    if k.ndim != nd:
        raise NdConvError(ERR_BAD_ARG, "data / kernel rank mismatch")
    if k.dtype != dtype:
        k = k.astype(dtype)
    pr = _Problem()
    pr.dtype, pr.ndim, pr.memory, pr.reverse = DTYPE_CODES[dtype], nd, memory, int(bool(kernel.reverse_flag))
    pr.data = x_ptr
    pr.kernel = k.ctypes.data
    if explicit is None:
        pads, strides = conv_mode.unfold(k.shape, kernel.dilation, lib)
    else:
        pads, strides = np.asarray(explicit[0], np.int64).reshape(nd, 2), np.asarray(explicit[1], np.int64)
    borders = padding_mode.lower(nd)
    for i in range(nd):
        pr.data_shape[i] = x_shape[i]
        pr.data_strides[i] = x_strides_elems[i]
        pr.kernel_shape[i] = k.shape[i]
        pr.kernel_strides[i] = k.strides[i] // k.itemsize
        pr.dilation[i] = kernel.dilation[i]
        pr.pad[i][0], pr.pad[i][1] = int(pads[i][0]), int(pads[i][1])
        pr.stride[i] = int(strides[i])
        for s in range(2):
            b = borders[i][s]
            pr.border[i][s].type = b.kind
            if b.kind == BorderType.CONST:
                raw = _scalar_bytes(b.value, dtype)
                for j, byte in enumerate(raw):
                    pr.border[i][s].value[j] = byte
    return pr, (k,)


def out_shape(pr, path, lib: Library | None = None):
    lib = lib or get_library()
    shp = (ctypes.c_int64 * MAX_DIM)()
    lib.check(lib.c.ndconv_out_shape(ctypes.byref(pr), path, shp))
    return [int(shp[i]) for i in range(pr.ndim)]


class Processor:
    """`impl Processor<T, InElem>` (src/conv_fft/processor/mod.rs:79-119) as a handle: owns the device, a
    stream, FFT plans, twiddles, the cached kernel spectrum and workspaces."""

    def __init__(self, device=0, lib: Library | None = None):
        self.lib = lib or get_library()
        h = ctypes.c_void_p()
        self.lib.check(self.lib.c.ndconv_processor_create(device, ctypes.byref(h)))
        self.handle = h

    def set_stream(self, cuda_stream_ptr):
        self.lib.check(self.lib.c.ndconv_processor_set_stream(self.handle, ctypes.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self.lib.check(self.lib.c.ndconv_processor_synchronize(self.handle))

    @property
    def launch_count(self):
        return int(self.lib.c.ndconv_processor_launch_count(self.handle))

    @property
    def workspace_bytes(self):
        return int(self.lib.c.ndconv_processor_workspace_bytes(self.handle))

    # Processor::forward / backward (src/conv_fft/processor/mod.rs:91-118) with the reference's rotated spectrum layout
    def forward(self, x):
        """N-d FFT of a real (f32/f64) or complex array; returns the spectrum with axis 0 moved to the end."""
        x = np.ascontiguousarray(x)
        cplx = np.iscomplexobj(x)
        cdt = np.result_type(x.dtype, np.complex64)
        last = x.shape[-1] if cplx else x.shape[-1] // 2 + 1
        oshape = (list(x.shape[1:-1]) + [last, x.shape[0]]) if x.ndim > 1 else [last]
        out = np.empty(oshape, cdt)
        shp = (ctypes.c_int64 * x.ndim)(*x.shape)
        self.lib.check(self.lib.c.ndconv_fft_forward(self.handle, DTYPE_CODES[x.dtype], x.ndim, shp, x.ctypes.data, out.ctypes.data, MEM_HOST))
        self._origin = (x.shape, x.dtype)       # the reference's rp_origin_len: backward needs the real-space last-axis length
        return out

    def backward(self, spectrum, shape=None, dtype=None):
        """inverse of forward (divides by the number of elements); shape / dtype default to those of the last forward()."""
        if shape is None:
            shape, dtype = self._origin
        dtype = np.dtype(dtype)
        spectrum = np.ascontiguousarray(spectrum, dtype=np.result_type(dtype, np.complex64))
        out = np.empty(shape, dtype)
        shp = (ctypes.c_int64 * len(shape))(*shape)
        self.lib.check(self.lib.c.ndconv_fft_backward(self.handle, DTYPE_CODES[dtype], len(shape), shp, spectrum.ctypes.data, out.ctypes.data, MEM_HOST))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.c.ndconv_processor_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_fft_processor(device=0, lib: Library | None = None) -> Processor:
    """get_fft_processor::<T, InElem>() (src/lib.rs:74-76, src/conv_fft/processor/mod.rs:71-73)"""
    return Processor(device, lib)


def _run_host(entry, path, x, kernel, conv_mode, padding_mode, processor, lib):
    lib = lib or (processor.lib if processor is not None else get_library())
    x = np.asarray(x)
    kwd = _into_kwd(kernel)
    pr, keep = make_problem(x.shape, [s // x.itemsize for s in x.strides], x.ctypes.data, x.dtype, kwd, conv_mode, padding_mode, MEM_HOST, lib)
    shp = out_shape(pr, path, lib)
    out = np.empty(shp, x.dtype)
    h = processor.handle if processor is not None else None
    lib.check(getattr(lib.c, entry)(h, ctypes.byref(pr), out.ctypes.data))
    del keep
    return out


def conv(x, kernel, conv_mode=ConvMode.Same, padding_mode=PaddingMode.Zeros, processor: Processor | None = None, lib=None):
    """ConvExt::conv (src/conv/mod.rs:110-115): direct convolution, any supported element type."""
    return _run_host("ndconv_conv_direct", PATH_DIRECT, x, kernel, conv_mode, padding_mode, processor, lib)


def conv_fft(x, kernel, conv_mode=ConvMode.Same, padding_mode=PaddingMode.Zeros, lib=None):
    """ConvFFTExt::conv_fft (src/conv_fft/mod.rs:394-402): fresh processor per call."""
    return _run_host("ndconv_conv_fft", PATH_FFT, x, kernel, conv_mode, padding_mode, None, lib)


def conv_fft_with_processor(x, kernel, conv_mode, padding_mode, processor: Processor):
    """ConvFFTExt::conv_fft_with_processor (src/conv_fft/mod.rs:404-412)"""
    return _run_host("ndconv_conv_fft", PATH_FFT, x, kernel, conv_mode, padding_mode, processor, None)


def conv_fft_par(x, kernel, conv_mode=ConvMode.Same, padding_mode=PaddingMode.Zeros, lib=None, processors=None):
    """ConvFFTExt::conv_fft_par (src/conv_fft/mod.rs:414-423): same GPU call -- the parallelism is the device's; with several
    `processors` configured (one per GPU) the one convolution is sharded over them (conv_fft_sharded)."""
    if processors is not None and len(processors) > 1:
        return conv_fft_sharded(x, kernel, conv_mode, padding_mode, processors)
    return _run_host("ndconv_conv_fft_par", PATH_FFT, x, kernel, conv_mode, padding_mode, processors[0] if processors else None, lib)


def conv_fft_sharded(x, kernel, conv_mode, padding_mode, processors, out=None):
    """conv_fft_par over several GPUs of this process (ndconv_conv_fft_sharded): one axis-0 slab of output rows per processor,
    each on its own host thread and PCIe link.  `x` / `out` should be pinned (ndconv_host_alloc) for full-rate copies."""
    lib = processors[0].lib
    x = np.asarray(x)
    kwd = _into_kwd(kernel)
    pr, keep = make_problem(x.shape, [s // x.itemsize for s in x.strides], x.ctypes.data, x.dtype, kwd, conv_mode, padding_mode, MEM_HOST, lib)
    shp = out_shape(pr, PATH_FFT, lib)
    if out is None:
        out = np.empty(shp, x.dtype)
    assert tuple(out.shape) == tuple(shp) and out.dtype == x.dtype and out.flags.c_contiguous
    handles = (ctypes.c_void_p * len(processors))(*[p.handle for p in processors])
    lib.check(lib.c.ndconv_conv_fft_sharded(handles, len(processors), ctypes.byref(pr), out.ctypes.data))
    del keep
    return out


def _global_problem(x_shape, dtype, kernel, conv_mode, padding_mode, lib):
    strides = [int(np.prod(x_shape[i + 1:])) for i in range(len(x_shape))]
    dummy = np.zeros(1, dtype)                        # the data pointer of the global problem is never dereferenced, but must not be null
    pr, keep = make_problem(x_shape, strides, dummy.ctypes.data, dtype, _into_kwd(kernel), conv_mode, padding_mode, MEM_DEVICE, lib)
    return pr, keep + (dummy,)


def shard_plan(x_shape, dtype, kernel, conv_mode, padding_mode, shard_rows, shard, lib=None):
    """ndconv_shard_plan: output rows shard `shard` produces and the ghost rows ndconv_conv_fft_sharded_device fills around its
    owned rows, for an array partitioned along axis 0 into len(shard_rows) device-resident shards (host logic only)."""
    lib = lib or get_library()
    pr, keep = _global_problem(x_shape, dtype, kernel, conv_mode, padding_mode, lib)
    rows = (ctypes.c_int64 * len(shard_rows))(*[int(r) for r in shard_rows])
    info = _ShardInfo()
    lib.check(lib.c.ndconv_shard_plan(ctypes.byref(pr), len(shard_rows), rows, shard, ctypes.byref(info)))
    return {k: int(getattr(info, k)) for k, _ in _ShardInfo._fields_}


def conv_fft_sharded_device(processors, x_shape, dtype, kernel, conv_mode, padding_mode, shards):
    """ndconv_conv_fft_sharded_device: one conv_fft over an array that is already device-resident and row-partitioned.
    shards[g] = dict(data=<device pointer of the first owned row>, rows=, halo_front=, halo_back=, out=<device pointer>); ghost rows
    are filled peer-to-peer, the pipeline runs in place on every shard.  Enqueued only: synchronise the processors afterwards."""
    lib = processors[0].lib
    pr, keep = _global_problem(x_shape, dtype, kernel, conv_mode, padding_mode, lib)
    arr = (_Shard * len(shards))()
    for i, sh in enumerate(shards):
        arr[i].data, arr[i].rows, arr[i].halo_front, arr[i].halo_back, arr[i].out = sh["data"], sh["rows"], sh["halo_front"], sh["halo_back"], sh["out"]
    handles = (ctypes.c_void_p * len(processors))(*[p.handle for p in processors])
    lib.check(lib.c.ndconv_conv_fft_sharded_device(handles, len(processors), ctypes.byref(pr), arr))
    del keep


class pinned:
    """context manager: page-lock numpy arrays the caller already owns for the duration of a block of host-resident calls
    (ndconv_host_register / ndconv_host_unregister)."""

    def __init__(self, *arrays, lib: Library | None = None):
        self.lib, self.arrays, self.done = lib or get_library(), arrays, []

    def __enter__(self):
        for a in self.arrays:
            self.lib.check(self.lib.c.ndconv_host_register(a.ctypes.data, a.nbytes))
            self.done.append(a)
        return self

    def __exit__(self, *exc):
        for a in self.done:
            self.lib.c.ndconv_host_unregister(a.ctypes.data)
        self.done = []
        return False


def conv_fft_batch(xs, kernels, conv_mode, padding_mode, processors):
    """independent conv_fft problems distributed whole over processor handles (ndconv_conv_fft_batch): xs[i] * kernels[i] (one
    kernel for all when `kernels` is not a list) runs on processors[i % len(processors)], each handle on its own host thread."""
    lib = processors[0].lib
    xs = [np.asarray(x) for x in xs]
    ks = kernels if isinstance(kernels, (list, tuple)) else [kernels] * len(xs)
    prs, keeps, outs = (_Problem * len(xs))(), [], []
    for i, (x, k) in enumerate(zip(xs, ks)):
        pr, keep = make_problem(x.shape, [s // x.itemsize for s in x.strides], x.ctypes.data, x.dtype, _into_kwd(k), conv_mode, padding_mode, MEM_HOST, lib)
        prs[i] = pr
        keeps.append(keep)
        outs.append(np.empty(out_shape(pr, PATH_FFT, lib), x.dtype))
    handles = (ctypes.c_void_p * len(processors))(*[p.handle for p in processors])
    optr = (ctypes.c_void_p * len(xs))(*[o.ctypes.data for o in outs])
    lib.check(lib.c.ndconv_conv_fft_batch(handles, len(processors), prs, optr, len(xs)))
    del keeps
    return outs


def conv_device(entry, processor: Processor, x_ptr, x_shape, x_strides_elems, dtype, kernel, conv_mode, padding_mode, out_ptr, explicit=None):
    """Device-resident call: x_ptr / out_ptr are device pointers (e.g. torch tensor .data_ptr()); the work is
    enqueued on the processor's stream and NOT synchronised.  Returns the output shape."""
    lib = processor.lib
    kwd = _into_kwd(kernel)
    pr, keep = make_problem(x_shape, x_strides_elems, x_ptr, dtype, kwd, conv_mode, padding_mode, MEM_DEVICE, lib, explicit=explicit)
    path = PATH_DIRECT if entry == "ndconv_conv_direct" else PATH_FFT
    shp = out_shape(pr, path, lib)
    if out_ptr is not None:
        lib.check(getattr(lib.c, entry)(processor.handle, ctypes.byref(pr), ctypes.c_void_p(out_ptr)))
    del keep
    return shp


class PreparedConv:
    """A lowered problem kept alive for repeated device-resident calls (only the data / output pointers change):
    the per-call host cost is one ctypes call; the processor's plan cache makes the C side a key lookup + launches."""

    def __init__(self, entry, processor: Processor, x_shape, x_strides_elems, dtype, kernel, conv_mode, padding_mode, explicit=None):
        self.lib = processor.lib
        self.processor = processor
        self.fn = getattr(self.lib.c, entry)
        kwd = _into_kwd(kernel)
        self.pr, self._keep = make_problem(x_shape, x_strides_elems, 0, dtype, kwd, conv_mode, padding_mode, MEM_DEVICE, self.lib, explicit=explicit)
        self.pr.data = 1   # placeholder so shape validation does not trip on a null pointer
        self.out_shape = out_shape(self.pr, PATH_DIRECT if entry == "ndconv_conv_direct" else PATH_FFT, self.lib)
        self._ref = ctypes.byref(self.pr)

    def __call__(self, x_ptr, out_ptr):
        self.pr.data = x_ptr
        st = self.fn(self.processor.handle, self._ref, ctypes.c_void_p(out_ptr))
        if st:
            self.lib.check(st)


def plan_query(x_shape, dtype, kernel, conv_mode, padding_mode, memory=MEM_DEVICE, lib=None):
    """what conv_fft would do with this problem (ndconv_plan_query; host logic only): dict(path, tile_len, tile_valid, n_tiles,
    workspace_bytes, split_out_rows, pipelined)"""
    lib = lib or get_library()
    kwd = _into_kwd(kernel)
    strides = [int(np.prod(x_shape[i + 1:])) for i in range(len(x_shape))]
    pr, keep = make_problem(tuple(x_shape), strides, 1, dtype, kwd, conv_mode, padding_mode, memory, lib)     # data pointer unused by the planner
    info = _PlanInfo()
    lib.check(lib.c.ndconv_plan_query(ctypes.byref(pr), ctypes.byref(info)))
    nd = info.ndim
    return {"path": ("generic", "fast", "direct", "split")[info.path], "tile_len": list(info.tile_len[:nd]), "tile_valid": list(info.tile_valid[:nd]),
            "n_tiles": list(info.n_tiles[:nd]), "workspace_bytes": int(info.workspace_bytes), "split_out_rows": int(info.split_out_rows),
            "pipelined": bool(info.pipelined)}


def slab_plan(x_shape, dtype, kernel, conv_mode, padding_mode, path, n_slabs, slab, lib=None):
    """ndconv_slab_plan: the output rows of axis 0 owned by `slab` and the padded rows it reads."""
    lib = lib or get_library()
    kwd = _into_kwd(kernel)
    dummy = np.zeros(1, dtype)
    strides = [int(np.prod(x_shape[i + 1:])) for i in range(len(x_shape))]
    pr, keep = make_problem(x_shape, strides, dummy.ctypes.data, dtype, kwd, conv_mode, padding_mode, MEM_HOST, lib)
    s = _Slab()
    lib.check(lib.c.ndconv_slab_plan(ctypes.byref(pr), path, n_slabs, slab, ctypes.byref(s)))
    return dict(out_begin=s.out_begin, out_end=s.out_end, pad_begin=s.pad_begin, pad_end=s.pad_end)
