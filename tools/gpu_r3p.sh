#!/bin/bash
# round 2, call 3P (1 GPU): compute-sanitizer over the round-2f kernels (twiddle_store, keep_range store epilogues, register-store variant of the Tensor-Memory column pass)
mkdir -p gpurun_out
export NDCONV_BLOCKED_MIN_OUT=0
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_cases.py 2>&1 | grep -vE "^\s*$" | tail -6
done | tee gpurun_out/r02f_sanitizer.txt
echo "== memcheck, NDCONV_COL_STG=1" | tee -a gpurun_out/r02f_sanitizer.txt
NDCONV_COL_STG=1 timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_cases.py 2>&1 | grep -vE "^\s*$" | tail -4 | tee -a gpurun_out/r02f_sanitizer.txt
