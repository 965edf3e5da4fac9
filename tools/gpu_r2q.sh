#!/bin/bash
# round 2, call Q (1 GPU): where row_fwd requests the next row into L2 (DRAM read traffic went 4.75 -> 5.45 GB with the early prefetch)
mkdir -p gpurun_out
for m in 0 1 2; do
  NDCONV_ROW_PF=$m ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k "regex:row_fwd" -s 2 -c 1 --csv \
    --log-file gpurun_out/r2q_pf$m.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e > /dev/null 2>&1
  echo "pf_mode $m"; tail -2 gpurun_out/r2q_pf$m.csv | cut -d, -f5,13-
done
tools/gpu_variants.sh "NDCONV_ROW_PF=0" "NDCONV_ROW_PF=1" "NDCONV_ROW_PF=2" "NDCONV_ROW_PF=0" "NDCONV_ROW_PF=1" "NDCONV_ROW_PF=2" 2>&1 | tee gpurun_out/r2q_variants.txt
