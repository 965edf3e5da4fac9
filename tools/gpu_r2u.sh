#!/bin/bash
# round 2, call U (1 GPU): register-blocked direct conv with vector loads: parity in a child with the size gate lowered, timings against the unblocked kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_direct_blocked.py tests/test_baseline_configs.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2u_tests.txt
python tools/run_direct_shapes.py 2>&1 | tee gpurun_out/r2u_direct_shapes.txt
NDCONV_DISABLE_BLOCKED=1 python tools/run_direct_shapes.py 2>&1 | tee -a gpurun_out/r2u_direct_shapes.txt
