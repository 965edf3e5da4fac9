#!/bin/bash
# round 2, call V (1 GPU): compute-sanitizer over every kernel family incl. the round-2b kernels (memcheck, racecheck, synccheck)
mkdir -p gpurun_out
export NDCONV_BLOCKED_MIN_OUT=0
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_cases.py 2>&1 | grep -vE "^\s*$" | tail -12
done | tee gpurun_out/r02b_sanitizer.txt
