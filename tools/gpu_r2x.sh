#!/bin/bash
# round 2, call X (1 GPU): persistent double-buffered direct conv: parity, timings against the one-tile-per-CTA blocked kernel and the tap-list kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_direct_blocked.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2x_tests.txt
python tools/run_direct_shapes.py 2>&1 | tee gpurun_out/r2x_direct_shapes.txt
NDCONV_DISABLE_PERSIST=1 python tools/run_direct_shapes.py 2>&1 | sed 's/blocked /blocked, one tile per CTA /' | tee -a gpurun_out/r2x_direct_shapes.txt
