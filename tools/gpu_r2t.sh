#!/bin/bash
# round 2, call T (1 GPU): register-blocked direct conv: parity (direct-conv tests), timings and bit-for-bit hashes against the unblocked kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_small.py tests/test_baseline_configs.py tests/test_int128.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2t_tests.txt
python tools/run_direct_shapes.py 2>&1 | tee gpurun_out/r2t_direct_shapes.txt
NDCONV_DISABLE_BLOCKED=1 python tools/run_direct_shapes.py 2>&1 | tee -a gpurun_out/r2t_direct_shapes.txt
