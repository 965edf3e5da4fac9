#!/bin/bash
# round 2, call B: col_pass_tma -- parity, racecheck/memcheck on small cases, A/B against col_pass on c5
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_parity_opt.py tests/test_baseline_configs.py -m gpu -x -q -k "not full_bands" ) 2>&1 | tail -6
echo "== sanitizer (memcheck + racecheck) on the 1024-row cases"
cat > /tmp/san.py <<PY
import importlib, sys, numpy as np
sys.path.insert(0, '.')
pkg = importlib.import_module("ndarray-conv_b200")
rng = np.random.default_rng(0)
proc = pkg.get_fft_processor(0)
k = rng.random((5, 7), dtype=np.float32)
pkg.conv_fft_with_processor(rng.random((1100, 2100), dtype=np.float32), k, pkg.ConvMode.Full, pkg.PaddingMode.Reflect, proc)
pkg.conv_fft_with_processor(rng.random((3, 1000, 300), dtype=np.float32), rng.random((2, 3, 5), dtype=np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Reflect, proc)
proc.close()
print("SAN_DONE")
PY
timeout 900 compute-sanitizer --tool memcheck python /tmp/san.py 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck python /tmp/san.py 2>&1 | tail -4
echo "== A/B on c5"
bash tools/gpu_variants.sh "NDCONV_DISABLE_COL_TMA=1" "X=1" "NDCONV_DISABLE_COL_TMA=1" "X=1"
