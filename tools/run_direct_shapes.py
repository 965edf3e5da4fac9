"""per-call times of direct conv shapes, blocked vs NDCONV_DISABLE_BLOCKED=1, and a bit-for-bit comparison of the two: python tools/run_direct_shapes.py"""
import importlib, sys, numpy as np, torch, os, hashlib
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
proc = pkg.get_fft_processor(0); dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); proc.set_stream(st.cuda_stream)
B = pkg.BorderType
CASES = [
    ("c4 i32 (64,256,256) k(3,5,5) s2 Replicate", np.int32, (64, 256, 256), (3, 5, 5), 1, pkg.ConvMode.Custom([1, 2, 2], [2, 2, 2]), pkg.PaddingMode.Replicate),
    ("i32 (256,1024,1024) k(3,5,5) Same Replicate", np.int32, (256, 1024, 1024), (3, 5, 5), 1, pkg.ConvMode.Same, pkg.PaddingMode.Replicate),
    ("f32 (256,1024,1024) k(3,3,3) Same Zeros", np.float32, (256, 1024, 1024), (3, 3, 3), 1, pkg.ConvMode.Same, pkg.PaddingMode.Zeros),
    ("f32 (8192,8192) k(7,7) Same Reflect", np.float32, (8192, 8192), (7, 7), 1, pkg.ConvMode.Same, pkg.PaddingMode.Reflect),
    ("f64 (4096,4096) k(5,5) dil2 Same Reflect", np.float64, (4096, 4096), (5, 5), 2, pkg.ConvMode.Same, pkg.PaddingMode.Reflect),
    ("i64 (64,512,512) k(3,3,3) s(1,2,2) Circular", np.int64, (64, 512, 512), (3, 3, 3), 1, pkg.ConvMode.Custom([1, 1, 1], [1, 2, 2]), pkg.PaddingMode.Circular),
]
for ci, (name, dt, xs, ks, dil, mode, pm) in enumerate(CASES):
    rng = np.random.default_rng(1000 + ci)
    if np.dtype(dt).kind == "i":
        xh = rng.integers(-128, 128, size=xs).astype(dt); kh = rng.integers(-8, 8, size=ks).astype(dt)
    else:
        xh = (rng.random(xs) - 0.5).astype(dt); kh = (rng.random(ks) - 0.5).astype(dt)
    kh.flat[1] = 0                                       # a zero tap (dropped from the tap list; a mask bit in the blocked rows)
    xd = torch.from_numpy(xh).to(dev)
    strides = [int(np.prod(xs[i + 1:])) for i in range(len(xs))]
    prep = pkg.PreparedConv("ndconv_conv_direct", proc, xs, strides, dt, pkg.with_dilation(kh, dil), mode, pm)
    yd = torch.empty(prep.out_shape, dtype=getattr(torch, np.dtype(dt).name), device=dev)
    for _ in range(3): prep(xd.data_ptr(), yd.data_ptr())
    torch.cuda.synchronize()
    n = 20 if np.prod(xs) < 1e8 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): prep(xd.data_ptr(), yd.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / n
    nout = int(np.prod(prep.out_shape))
    gb = (xh.nbytes + nout * xh.itemsize) / 1e9
    h = hashlib.sha256(yd.cpu().numpy().tobytes()).hexdigest()[:16]
    print("%-48s %-8s %9.3f us/call %8.1f Gsamples/s %7.1f GB/s compulsory  sha %s" % (name, "unblocked" if os.environ.get("NDCONV_DISABLE_BLOCKED") else "blocked", per * 1e3, nout / per / 1e6, gb / per * 1e3, h), flush=True)
