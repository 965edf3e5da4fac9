import importlib, sys, numpy as np, subprocess, os
sys.path.insert(0, '.')
code = '''
import importlib, sys, numpy as np
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
dt = np.dtype(sys.argv[1]); shape = eval(sys.argv[2]); ks = eval(sys.argv[3])
x = np.arange(1, int(np.prod(shape)) + 1).reshape(shape).astype(dt)
k = np.ones(ks, dt)
y = pkg.conv(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Zeros)
print("ok", dt, shape, y.ravel()[:6])
'''
for dt, shape, ks in [("int32", (2, 2), (2, 2)), ("int64", (2, 2), (2, 2)), ("int32", (8, 8), (3, 3)), ("float32", (16, 64), (3, 3)), ("int64", (8, 8), (3, 3)), ("int32", (4, 8, 16), (1, 3, 3))]:
    r = subprocess.run([sys.executable, "-c", code, dt, str(shape), str(ks)], capture_output=True, text=True)
    print(dt, shape, ks, "->", (r.stdout.strip() or r.stderr.strip()[-300:]))
