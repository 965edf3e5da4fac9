"""c2-shaped conv_fft with different borders: how much of the call is the edge path of row_fwd?  (device-resident, warm processor)"""
import importlib, sys, numpy as np, torch, ctypes
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200"); lib = pkg.get_library()
proc = pkg.get_fft_processor(0); dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
x = torch.rand((200, 5000), device=dev)
k = np.random.default_rng(2001).random((11, 31), dtype=np.float32)
B = pkg.BorderType
cases = [("c2: Same, [Reflect, Circular]", pkg.ConvMode.Same, pkg.PaddingMode.Custom([B.Reflect, B.Circular])),
         ("Same, Zeros", pkg.ConvMode.Same, pkg.PaddingMode.Zeros),
         ("Same, Replicate", pkg.ConvMode.Same, pkg.PaddingMode.Replicate),
         ("Valid (no border)", pkg.ConvMode.Valid, pkg.PaddingMode.Zeros),
         ("Full, Reflect", pkg.ConvMode.Full, pkg.PaddingMode.Reflect)]
for name, mode, pm in cases:
    prep = pkg.PreparedConv("ndconv_conv_fft", proc, (200, 5000), (5000, 1), np.float32, pkg.with_dilation(k, 2), mode, pm)
    y = torch.empty(prep.out_shape, dtype=torch.float32, device=dev)
    for _ in range(20): prep(x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(200): prep(x.data_ptr(), y.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 200 * 1e3
    lib.c.ndconv_processor_set_profiling(proc.handle, 1)
    for _ in range(20): prep(x.data_ptr(), y.data_ptr())
    names = (ctypes.c_char * 64 * 16)(); ms = (ctypes.c_double * 16)(); cnt = (ctypes.c_int64 * 16)(); by = (ctypes.c_double * 16)()
    n = lib.c.ndconv_processor_get_profile(proc.handle, 16, names, ms, cnt, by)
    lib.c.ndconv_processor_set_profiling(proc.handle, 0)
    ker = {bytes(names[i]).split(b"\0")[0].decode(): round(ms[i] / cnt[i] * 1e3, 1) for i in range(n)}
    print("%-32s %6.1f us/call  out %s  per-kernel us (events, serialised) %s" % (name, per, tuple(prep.out_shape), ker), flush=True)
