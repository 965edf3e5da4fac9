#!/bin/bash
# ncu evidence for profiles/ (one GPU; numbers printed under ncu are never bench values).  usage: tools/gpu_profile.sh <tag>
tag=${1:-r01b}
K='regex:row_fwd|col_pass|row_inv'
# 1. launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c5_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e > gpurun_out/ncu_launch_$tag.log 2>&1
# 2. DRAM traffic of one launch of each fast-path kernel on the FULL c5 workload
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" -s 6 -c 3 --csv \
    --log-file gpurun_out/traffic_c5_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e > gpurun_out/ncu_traffic_$tag.log 2>&1
# 3. --set full of the three kernels on the reduced workload c5s (same kernels, same tile shape)
ncu --set full --clock-control none --import-source on -k "$K" -s 6 -c 3 -o gpurun_out/prof_c5s_$tag -f \
    python bench.py --workload c5s --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -8
