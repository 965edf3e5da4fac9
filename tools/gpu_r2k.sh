#!/bin/bash
# round 2, call K (1 GPU): compute-sanitizer over every kernel family (memcheck, racecheck, synccheck)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_cases.py 2>&1 | grep -vE "^\s*$" | tail -6
done | tee gpurun_out/r02_sanitizer.txt
