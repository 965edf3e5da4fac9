#!/bin/bash
# round 2, call 4F (1 GPU): blocked direct loop with the tail tap groups behind uniform branches: parity (bit-identical SHA against the tap-list kernel), per-shape times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_direct_blocked.py tests/test_int128.py tests/test_parity_small.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r4f_tests.txt
python tools/run_direct_shapes.py 2>&1 | cut -c1-200 | tee gpurun_out/r4f_direct_shapes.txt
