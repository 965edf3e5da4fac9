"""per-kernel times of f64 conv_fft shapes (fast path vs NDCONV_DISABLE_OPT64=1 generic kernels): python tools/run_f64_shapes.py"""
import importlib, sys, numpy as np, torch, ctypes, os
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200"); lib = pkg.get_library()
proc = pkg.get_fft_processor(0); dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
rng = np.random.default_rng(0)
for xs, ks, mode, pm in (((8192, 8192), (63, 63), pkg.ConvMode.Full, pkg.PaddingMode.Reflect), ((4096, 4096), (63, 63), pkg.ConvMode.Same, pkg.PaddingMode.Reflect),
                         ((200, 5000), (11, 31), pkg.ConvMode.Same, pkg.PaddingMode.Zeros), ((64, 256, 512), (5, 7, 9), pkg.ConvMode.Same, pkg.PaddingMode.Replicate)):
    x = torch.from_numpy(rng.random(xs)).to(dev); k = rng.random(ks)
    strides = [int(np.prod(xs[i + 1:])) for i in range(len(xs))]
    prep = pkg.PreparedConv("ndconv_conv_fft", proc, xs, strides, np.float64, k, mode, pm)
    y = torch.empty(prep.out_shape, dtype=torch.float64, device=dev)
    for _ in range(3): prep(x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): prep(x.data_ptr(), y.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 10
    lib.c.ndconv_processor_set_profiling(proc.handle, 1)
    for _ in range(5): prep(x.data_ptr(), y.data_ptr())
    names = (ctypes.c_char * 64 * 16)(); ms = (ctypes.c_double * 16)(); cnt = (ctypes.c_int64 * 16)(); by = (ctypes.c_double * 16)()
    n = lib.c.ndconv_processor_get_profile(proc.handle, 16, names, ms, cnt, by)
    lib.c.ndconv_processor_set_profiling(proc.handle, 0)
    ker = {bytes(names[i]).split(b"\0")[0].decode(): (round(ms[i] / cnt[i], 3), round(by[i] / cnt[i] / (ms[i] / cnt[i]) / 1e6, 1)) for i in range(n)}
    print("f64", xs, ks, "OPT64 off" if os.environ.get("NDCONV_DISABLE_OPT64") else "fast", round(per, 3), "ms/call", round(np.prod(prep.out_shape) / per / 1e6, 1), "Gsamples/s", ker, flush=True)
