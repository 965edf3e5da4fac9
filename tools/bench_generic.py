"""device-resident timings of problems that take the GENERIC kernels (f64, complex, 3-D) next to a fast-path case"""
import importlib, sys, numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
lib = pkg.get_library()
proc = pkg.get_fft_processor(0)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
cases = [("f32 4096^2 k63 (fast path)", np.float32, (4096, 4096), (63, 63)), ("f64 4096^2 k63", np.float64, (4096, 4096), (63, 63)),
         ("c64 4096^2 k63", np.complex64, (4096, 4096), (63, 63)), ("f32 256^3 k15", np.float32, (256, 256, 256), (15, 15, 15)),
         ("f32 1d 16M k1025", np.float32, (1 << 24,), (1025,)), ("f32 2-D 4096x1000 k31 (P1<1200: generic)", np.float32, (4096, 1000), (31, 31))]
for name, dt, xs, ks in cases:
    rng = np.random.default_rng(0)
    if dt == np.complex64:
        x = (rng.random(xs, dtype=np.float32) + 1j * rng.random(xs, dtype=np.float32)).astype(dt); k = (rng.random(ks, dtype=np.float32) + 0j).astype(dt)
        xt = torch.from_numpy(x.view(np.float32)).to(dev)
    else:
        x = rng.random(xs).astype(dt); k = rng.random(ks).astype(dt); xt = torch.from_numpy(x).to(dev)
    strides = [int(np.prod(xs[i + 1:])) for i in range(len(xs))]
    prep = pkg.PreparedConv("ndconv_conv_fft", proc, xs, strides, dt, k, pkg.ConvMode.Same, pkg.PaddingMode.Reflect)
    n_out = int(np.prod(prep.out_shape))
    y = torch.empty(n_out * (2 if dt == np.complex64 else 1), dtype=torch.float64 if dt == np.float64 else torch.float32, device=dev)
    for _ in range(3): prep(xt.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5): prep(xt.data_ptr(), y.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:45s} {ms:9.3f} ms  {n_out / ms / 1e6:8.2f} Gsamples/s")
