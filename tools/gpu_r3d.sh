#!/bin/bash
# round 2, call 3D (1 GPU): stress (fresh processors over poisoned memory, repeated) incl. the Tensor-Memory column pass; compute-sanitizer over the final kernels
mkdir -p gpurun_out
timeout 1200 python tools/stress_case.py 8 2>&1 | tail -6 | tee gpurun_out/r3d_stress.txt
export NDCONV_BLOCKED_MIN_OUT=0
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_cases.py 2>&1 | grep -vE "^\s*$" | tail -8
done | tee gpurun_out/r02c_sanitizer.txt
export NDCONV_PERSIST_MIN_TILES=0 NDCONV_PERSIST_MAX_GRID=3
echo "== racecheck, persistent direct variant forced" | tee -a gpurun_out/r02c_sanitizer.txt
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_cases.py 2>&1 | grep -vE "^\s*$" | tail -4 | tee -a gpurun_out/r02c_sanitizer.txt
