#!/bin/bash
# round 2, call 3J (1 GPU): ncu --set full with source of the three fast-path kernels on the FULL c5 workload; per-instruction stall samples
mkdir -p gpurun_out /tmp/rep
K='regex:row_fwd|col_pass|row_inv'
ncu --set full --clock-control none --import-source on -k "$K" -s 6 -c 3 -o /tmp/rep/c5full -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e --no-pageable > gpurun_out/ncu_c5full_r02e.log 2>&1
python tools/summarize_ncu.py /tmp/rep/c5full.ncu-rep gpurun_out/r02e_ncu_full_c5 | cut -c1-600
ncu -i /tmp/rep/c5full.ncu-rep --page source --csv > gpurun_out/r02e_c5full_source.csv 2>/dev/null
ls -la /tmp/rep gpurun_out/r02e_c5full_source.csv
cp /tmp/rep/c5full.ncu-rep gpurun_out/ 2>/dev/null
