#!/bin/bash
# round 2, call 3X (1 GPU): final state -- the whole GPU suite, smoke, stress over poisoned workspaces, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3x_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r3x_smoke.txt
timeout 900 python tools/stress_case.py 4 2>&1 | tail -4 | tee gpurun_out/r3x_stress.txt
python bench.py > gpurun_out/r3x_bench.json 2> gpurun_out/r3x_bench.err; tail -c 400 gpurun_out/r3x_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3x_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],2), "e2e ms", round(d["e2e"]["ms_per_step"],1), "floor frac", d["e2e"].get("frac_of_copy_floor"))
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "pipeline", round(d["pipeline_compulsory"]["frac"],3), "cpu", d.get("cpu_baseline",{}).get("value"), "pageable", (d.get("e2e_pageable") or {}).get("value"))
print([(k["kernel"][:14], round(k["avg_ms"],3), round(k["frac_of_peak"],3)) for k in d["kernels"]])
for s in d["other_shapes"]: print(s["shape"][:70], round(s["us_per_call"],1), "us")
PY
