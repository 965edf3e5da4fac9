#!/bin/bash
# round 2, call 3V (1 GPU): the column pass's transform chain alone (experimental build, workspace loads and stores off) under ncu --set full with source
mkdir -p gpurun_out /tmp/rep
cp tools/exp/bin/libndconv_cuda_ring.so ndarray-conv_b200/libndconv_cuda.so
NDCONV_EXP_FLAGS=3 ncu --set full --clock-control none --import-source on -k "regex:col_pass_tma_kres" -s 2 -c 1 -o /tmp/rep/colchain -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e --no-pageable > gpurun_out/ncu_colchain_r02f.log 2>&1
python tools/summarize_ncu.py /tmp/rep/colchain.ncu-rep gpurun_out/r02f_ncu_full_colchain | cut -c1-600
ncu -i /tmp/rep/colchain.ncu-rep --page source --csv > gpurun_out/r02f_colchain_source.csv 2>/dev/null
ls -la gpurun_out/r02f_colchain_source.csv
