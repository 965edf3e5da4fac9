#!/bin/bash
# round 2, call E (1 GPU): col_pass_tma after the drain-race fix -- tests, poison stress, A/B of the store variants
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_int128.py tests/test_sharded_device.py tests/test_parity_opt.py tests/test_baseline_configs.py -m gpu -x -q -k "not full_bands" ) 2>&1 | tail -5
echo "== stress, TMA column kernel"
timeout 600 python tools/stress_case.py 12 2>&1 | tail -6
echo "== A/B on c5"
bash tools/gpu_variants.sh "NDCONV_DISABLE_COL_TMA=1" "X=1" "NDCONV_COL_STG=1" "X=1" "NDCONV_COL_STG=1"
