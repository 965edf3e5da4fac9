#!/bin/bash
# round 2, call W (1 GPU): the register-blocked direct kernel on the small c4 (size gate lowered) against the tap-list kernel
for g in 0 100000000; do
  echo "NDCONV_BLOCKED_MIN_OUT=$g"; NDCONV_BLOCKED_MIN_OUT=$g NDCONV_DEBUG_BLOCKED=1 python tools/run_direct_shapes.py 2>&1 | grep -v "^\[ndconv\]" | head -3
done | tee gpurun_out/r2w_c4_blocked.txt
