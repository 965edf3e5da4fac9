"""run BASELINE configs c1-c4 a few times each (for an ncu launch list of the small shapes): python tools/run_shapes.py"""
import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import bench  # noqa: E402

pkg = importlib.import_module("ndarray-conv_b200")
lib = pkg.get_library()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
proc = pkg.get_fft_processor(0, lib)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
proc.set_stream(stream.cuda_stream)
for s in bench.other_shapes(pkg, lib, proc, dev, stream):
    print(s["shape"], round(s["us_per_call"], 1), "us/call", s["kernel_us_per_call"])
proc.close()
