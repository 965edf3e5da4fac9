import importlib, sys, numpy as np
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
dt = np.dtype(sys.argv[1]); shape = eval(sys.argv[2]); ks = eval(sys.argv[3])
x = np.arange(1, int(np.prod(shape)) + 1).reshape(shape).astype(dt)
k = np.ones(ks, dt)
y = pkg.conv(x, k, pkg.ConvMode.Full, pkg.PaddingMode.Zeros)
print("ok", dt, shape, y.ravel()[:6])
