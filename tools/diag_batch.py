import importlib, sys, os
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import oracle
pkg = importlib.import_module("ndarray-conv_b200")
rng = np.random.default_rng(9)
B = pkg.BorderType
x = rng.random((6, 200, 5000), dtype=np.float32) - 0.4
k = rng.random((1, 11, 31), dtype=np.float32) - 0.5
proc = pkg.get_fft_processor(0)
pm3 = pkg.PaddingMode.Custom([B.Zeros, B.Reflect, B.Circular]); pm2 = pkg.PaddingMode.Custom([B.Reflect, B.Circular])
got = pkg.conv_fft_with_processor(x, pkg.with_dilation(k, [1, 2, 2]), pkg.ConvMode.Same, pm3, proc)
each = np.stack([pkg.conv_fft_with_processor(x[b], pkg.with_dilation(k[0], 2), pkg.ConvMode.Same, pm2, proc) for b in range(6)])
print("plan folded:", pkg.plan_query((200, 5000), np.float32, pkg.with_dilation(k[0], 2), pkg.ConvMode.Same, pm2))
for b in range(6):
    ref = oracle.conv_f64_truth(x[b], k[0], "same", ("custom", ["reflect", "circular"]), 2, True)
    d = np.abs(got[b] - each[b]); e1 = np.abs(got[b] - ref); e2 = np.abs(each[b] - ref)
    w = np.argwhere(d > 1e-3)
    print(b, "folded-vs-each max", d.max(), "folded err", e1.max(), "each err", e2.max(), "max|ref|", np.abs(ref).max(), "bad", len(w), w[:2].tolist(), w[-2:].tolist())
