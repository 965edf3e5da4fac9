"""c2 (BASELINE configs[1]) conv_fft, device-resident, a few calls -- used under ncu."""
import importlib, sys, numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
proc = pkg.get_fft_processor(0)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
x = torch.rand((200, 5000), device=dev)
k = np.random.default_rng(2001).random((11, 31), dtype=np.float32)
pm = pkg.PaddingMode.Custom([pkg.BorderType.Reflect, pkg.BorderType.Circular])
prep = pkg.PreparedConv("ndconv_conv_fft", proc, (200, 5000), (5000, 1), np.float32, pkg.with_dilation(k, 2), pkg.ConvMode.Same, pm)
y = torch.empty(prep.out_shape, dtype=torch.float32, device=dev)
for _ in range(6):
    prep(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
print("done", prep.out_shape, pkg.plan_query((200, 5000), np.float32, pkg.with_dilation(k, 2), pkg.ConvMode.Same, pm))
