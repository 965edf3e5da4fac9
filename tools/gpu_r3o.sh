#!/bin/bash
# round 2, call 3O (1 GPU): final state of round 2f -- the whole GPU suite, smoke, the default bench line, the ncu launch list and DRAM traffic of the same command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3o_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r3o_smoke.txt
python bench.py > gpurun_out/r3o_bench.json 2> gpurun_out/r3o_bench.err; tail -c 400 gpurun_out/r3o_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3o_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],2), "e2e ms", round(d["e2e"]["ms_per_step"],1), "floor frac", d["e2e"].get("frac_of_copy_floor"))
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "pipeline", round(d["pipeline_compulsory"]["frac"],3), "cpu", d.get("cpu_baseline",{}).get("value"), "pageable", (d.get("e2e_pageable") or {}).get("value"))
print([(k["kernel"][:14], round(k["avg_ms"],3), round(k["frac_of_peak"],3)) for k in d["kernels"]])
for s in d["other_shapes"]: print(s["shape"][:70], round(s["us_per_call"],1), "us")
PY
K='regex:row_fwd|col_pass|row_inv'
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c5_r02f.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e --no-pageable > gpurun_out/ncu_launch_r02f.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" -s 6 -c 3 --csv \
    --log-file gpurun_out/traffic_c5_r02f.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e --no-pageable > gpurun_out/ncu_traffic_r02f.log 2>&1
tail -4 gpurun_out/traffic_c5_r02f.csv | cut -c1-300
