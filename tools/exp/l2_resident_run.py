"""A conv_fft whose whole spectra workspace (4 tiles of 1024 x 2048 = 34 MB) fits the 126 MB L2, run a few times device-resident:
under `ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,...` the per-kernel DRAM bytes tell whether the
workspace written by one pass is still L2-resident when the next pass reads it (the premise of a fused per-tile pipeline)."""
import importlib, sys, numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
proc = pkg.get_fft_processor(0)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
shape = tuple(int(v) for v in sys.argv[1:3]) if len(sys.argv) > 2 else (1900, 3900)
x = torch.rand(shape, device=dev)
k = np.random.default_rng(1).random((63, 63), dtype=np.float32)
prep = pkg.PreparedConv("ndconv_conv_fft", proc, shape, (shape[1], 1), np.float32, pkg.with_dilation(k, 1), pkg.ConvMode.Full, pkg.PaddingMode.Reflect)
y = torch.empty(prep.out_shape, dtype=torch.float32, device=dev)
for _ in range(4):
    prep(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
print("done", prep.out_shape, "workspace MB", proc.workspace_bytes / 1e6)
