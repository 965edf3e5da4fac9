#!/bin/bash
# round 2, call N (1 GPU): col_pass_tma_kres (kernel spectrum resident in Tensor Memory): parity, then A/B on c5
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_col_kres.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2n_tests.txt
tools/gpu_variants.sh "NDCONV_DISABLE_COL_KRES=1" "NDCONV_X=1" "NDCONV_COL_KRES_BUNDLE=10" "NDCONV_COL_KRES_BUNDLE=16" "NDCONV_COL_KRES_BUNDLE=24" 2>&1 | tee gpurun_out/r2n_variants.txt
