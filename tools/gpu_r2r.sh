#!/bin/bash
# round 2, call R (1 GPU): f64 fast path: parity, the whole GPU suite (random sweeps now route f64 through it), timings against the generic kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_f64_fast.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2r_tests_f64.txt
python tools/run_f64_shapes.py 2>&1 | tee gpurun_out/r2r_f64_shapes.txt
NDCONV_DISABLE_OPT64=1 python tools/run_f64_shapes.py 2>&1 | tee -a gpurun_out/r2r_f64_shapes.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2r_tests_all.txt
