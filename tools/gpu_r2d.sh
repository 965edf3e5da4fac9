#!/bin/bash
# round 2, call D (1 GPU): new tests (int128, zero-tap NaN, sharded_device on one device), stress of the rank-3 1024-row cases
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_int128.py tests/test_sharded_device.py tests/test_parity_opt.py -m gpu -x -q -s ) 2>&1 | grep -v "^\[sharded" | tail -12
echo "== stress, TMA column kernel"
timeout 600 python tools/stress_case.py 10 2>&1 | tail -12
echo "== stress, NDCONV_DISABLE_COL_TMA=1"
NDCONV_DISABLE_COL_TMA=1 timeout 600 python tools/stress_case.py 10 2>&1 | tail -12
echo "== sharded_device on one GPU (2 handles, 16384^2)"
timeout 300 python tools/run_sharded_device.py 16384 1 2>&1 | tail -2
