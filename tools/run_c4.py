"""c4 (BASELINE configs[3]) direct conv, device-resident, a few calls -- used under ncu."""
import importlib, sys, numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
proc = pkg.get_fft_processor(0)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
rng = np.random.default_rng(1003)
x = torch.from_numpy(rng.integers(-128, 128, size=(64, 256, 256), dtype=np.int32)).to(dev)
k = np.random.default_rng(2003).integers(-128, 128, size=(3, 5, 5), dtype=np.int32)
prep = pkg.PreparedConv("ndconv_conv_direct", proc, (64, 256, 256), (65536, 256, 1), np.int32, k, pkg.ConvMode.Custom([1, 2, 2], [2, 2, 2]), pkg.PaddingMode.Replicate)
y = torch.empty(prep.out_shape, dtype=torch.int32, device=dev)
for _ in range(6):
    prep(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
print("done", prep.out_shape)
