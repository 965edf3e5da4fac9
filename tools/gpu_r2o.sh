#!/bin/bash
# round 2, call O (1 GPU): row_fwd (edge tiles through the vector path, next row resolved + L2-prefetched early) and twiddles in Tensor Memory
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_col_kres.py tests/test_parity_opt.py tests/test_baseline_configs.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2o_tests.txt
tools/gpu_variants.sh "NDCONV_COL_KRES_NO_TWT=1" "NDCONV_X=1" "NDCONV_COL_KRES_NO_TWT=1" "NDCONV_X=1" 2>&1 | tee gpurun_out/r2o_variants.txt
