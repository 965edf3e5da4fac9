"""one large register-blocked direct conv (3-D i32) and one persistent rank-2 f32 call (used under ncu)"""
import importlib, sys, numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
proc = pkg.get_fft_processor(0); dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
rng = np.random.default_rng(0)
for dt, xs, ks, mode, pm in ((np.int32, (128, 1024, 1024), (3, 5, 5), pkg.ConvMode.Same, pkg.PaddingMode.Replicate), (np.float32, (8192, 8192), (7, 7), pkg.ConvMode.Same, pkg.PaddingMode.Reflect)):
    xh = rng.integers(-9, 9, size=xs).astype(dt) if np.dtype(dt).kind == "i" else rng.random(xs).astype(dt)
    kh = rng.integers(-3, 3, size=ks).astype(dt) if np.dtype(dt).kind == "i" else rng.random(ks).astype(dt)
    x = torch.from_numpy(xh).to(dev)
    strides = [int(np.prod(xs[i + 1:])) for i in range(len(xs))]
    prep = pkg.PreparedConv("ndconv_conv_direct", proc, xs, strides, dt, kh, mode, pm)
    y = torch.empty(prep.out_shape, dtype=getattr(torch, np.dtype(dt).name), device=dev)
    for _ in range(3): prep(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize(); print("done")
