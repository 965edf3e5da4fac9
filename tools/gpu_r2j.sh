#!/bin/bash
# round 2, call J (1 GPU): row kernels in lockstep (A/B), long kernels + Bluestein + new tests on the device
mkdir -p gpurun_out
echo "== A/B lockstep rows on c5"
bash tools/gpu_variants.sh "X=1" "NDCONV_ROW_LOCKSTEP=1" "X=1" "NDCONV_ROW_LOCKSTEP=1"
echo "== lockstep parity"
NDCONV_ROW_LOCKSTEP=1 timeout 600 python -m pytest tests/test_parity_opt.py tests/test_baseline_configs.py -m gpu -x -q -k "not full_bands" 2>&1 | tail -3
echo "== new tests"
( time timeout 900 python -m pytest tests/test_long_kernels.py tests/test_processor_fft.py tests/test_batch_fold.py tests/test_int128.py -m gpu -x -q ) 2>&1 | tail -5
