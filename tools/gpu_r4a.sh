#!/bin/bash
# round 2, call 4A (1 GPU): c2-shaped calls with different borders
mkdir -p gpurun_out
python tools/run_c2_variants.py 2>&1 | tee gpurun_out/r4a_c2_variants.txt
