#!/bin/bash
# round 2, call 3B (1 GPU): ncu --set full of the f64 fast-path kernels and of the register-blocked / persistent direct kernels, summarised on the box
mkdir -p gpurun_out /tmp/rep
ncu --set full --clock-control none --import-source on -k "regex:row_fwd_d|col_pass_d|row_inv_d" -s 6 -c 3 -o /tmp/rep/prof_f64_r02c -f python tools/run_f64_one.py > gpurun_out/ncu_f64_r02c.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:direct_tile" -s 2 -c 1 -o /tmp/rep/prof_direct_blocked_r02d -f python tools/run_direct_one.py > gpurun_out/ncu_direct_r02d.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:direct_tile_persistent" -s 1 -c 1 -o /tmp/rep/prof_direct_persist_r02d -f python tools/run_direct_one.py > gpurun_out/ncu_direct2_r02d.log 2>&1
python tools/summarize_ncu.py /tmp/rep/prof_f64_r02c.ncu-rep gpurun_out/r02c_ncu_full_f64_8192 > /dev/null
python tools/summarize_ncu.py /tmp/rep/prof_direct_blocked_r02d.ncu-rep gpurun_out/r02d_ncu_full_direct_blocked > /dev/null
python tools/summarize_ncu.py /tmp/rep/prof_direct_persist_r02d.ncu-rep gpurun_out/r02d_ncu_full_direct_persistent > /dev/null
cat gpurun_out/r02c_ncu_full_f64_8192.md gpurun_out/r02d_ncu_full_direct_blocked.md gpurun_out/r02d_ncu_full_direct_persistent.md | cut -c1-700
