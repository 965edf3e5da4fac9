#!/bin/bash
# round 2, call C: the failing rank-3 case under both column kernels, then a full ncu capture of col_pass_tma on c5s
mkdir -p gpurun_out
for v in "X=1" "NDCONV_DISABLE_COL_TMA=1"; do
echo "== [$v]"
env $v timeout 300 python -m pytest tests/test_parity_opt.py -m gpu -q -k "1000" 2>&1 | grep -E "passed|failed|Error|assert|^E " | head -12
done
ncu --set full --clock-control none --import-source on -k "regex:col_pass" -s 2 -c 1 -o gpurun_out/prof_c5s_r02c_col -f \
    python bench.py --workload c5s --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e > gpurun_out/ncu_full_r02c.log 2>&1
ls -la gpurun_out | tail -4
