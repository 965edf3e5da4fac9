"""Host-resident conv_fft through the C ABI from ordinary (pageable) arrays against pinned ones, plus the cost of page-locking an
existing allocation (cudaHostRegister).  usage: python tools/pageable_probe.py [n=8192]"""
import ctypes
import importlib
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("ndarray-conv_b200")
lib = pkg.get_library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(0)
k = rng.random((63, 63), dtype=np.float32)
proc = pkg.get_fft_processor(0, lib)
x = rng.random((n, n), dtype=np.float32)
m = n + 62
out = np.zeros((m, m), np.float32)                  # touched: no first-touch page faults inside the timed calls


def run(xa, oa, tag):
    pr, keep = pkg.make_problem(xa.shape, (xa.shape[1], 1), xa.ctypes.data, np.float32, pkg._into_kwd(k), pkg.ConvMode.Full, pkg.PaddingMode.Reflect, pkg.MEM_HOST, lib)
    lib.check(lib.c.ndconv_conv_fft(proc.handle, ctypes.byref(pr), oa.ctypes.data))
    t0 = time.perf_counter()
    for _ in range(3):
        lib.check(lib.c.ndconv_conv_fft(proc.handle, ctypes.byref(pr), oa.ctypes.data))
    t = (time.perf_counter() - t0) / 3
    print(f"{tag}: {t * 1e3:.1f} ms per call, {(xa.nbytes + oa.nbytes) / t / 1e9:.1f} GB/s host<->device aggregate", flush=True)


run(x, out, "pageable")
cudart = ctypes.CDLL("libcudart.so")
for arr, name in ((x, "in"), (out, "out")):
    t0 = time.perf_counter()
    rc = cudart.cudaHostRegister(ctypes.c_void_p(arr.ctypes.data), ctypes.c_size_t(arr.nbytes), 0)
    print(f"cudaHostRegister({name}, {arr.nbytes / 1e6:.0f} MB) rc={rc}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
run(x, out, "registered in place")
for arr in (x, out):
    t0 = time.perf_counter()
    cudart.cudaHostUnregister(ctypes.c_void_p(arr.ctypes.data))
    print(f"cudaHostUnregister: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
proc.close()
