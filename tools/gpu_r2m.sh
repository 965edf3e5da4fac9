#!/bin/bash
# round 2, call M (1 GPU): TMEM probe (thread-private constant store), upper bounds for the column pass without kernel-spectrum / twiddle loads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 120 tools/exp/bin/tmem_probe 2>&1 | tee gpurun_out/r2m_tmem_probe.txt
tools/gpu_variants.sh "NDCONV_NONE=1" "NDCONV_COL_EXP=1" "NDCONV_COL_EXP=2" "NDCONV_COL_EXP=3" 2>&1 | tee gpurun_out/r2m_variants.txt
