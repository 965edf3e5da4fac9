#!/bin/bash
# round 2, call Y (1 GPU): direct conv after the map-staging reorder and the rank rule of the persistent variant: parity + timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_direct_blocked.py tests/test_parity_small.py tests/test_baseline_configs.py tests/test_int128.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2y_tests.txt
python tools/run_direct_shapes.py 2>&1 | tee gpurun_out/r2y_direct_shapes.txt
