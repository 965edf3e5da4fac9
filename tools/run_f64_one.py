"""one f64 fast-path shape, a few calls (used under ncu)"""
import importlib, sys, numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
proc = pkg.get_fft_processor(0); dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
rng = np.random.default_rng(0)
xs, ks = (8192, 8192), (63, 63)
x = torch.from_numpy(rng.random(xs)).to(dev); k = rng.random(ks)
prep = pkg.PreparedConv("ndconv_conv_fft", proc, xs, (8192, 1), np.float64, k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect)
y = torch.empty(prep.out_shape, dtype=torch.float64, device=dev)
for _ in range(4): prep(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize(); print("done")
