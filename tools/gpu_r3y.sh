#!/bin/bash
# round 2, call 3Y (1 GPU): repeatability of the e2e leg of the default bench line
mkdir -p gpurun_out
for i in 1 2 3; do
  python bench.py --no-shapes > gpurun_out/r3y.json 2> gpurun_out/r3y.err || tail -3 gpurun_out/r3y.err
  python - $i <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3y.json").read().strip().splitlines()[-1])
e=d["e2e"]; q=d.get("e2e_pageable") or {}
print("run", sys.argv[1], "| step %.3f ms" % d["ms_per_step"], "| e2e %.1f ms" % e["ms_per_step"], "floor %.1f ms" % (e.get("copy_floor_ms") or 0), "| pageable", round(q.get("ms_per_step", 0),1), "| clocks", d["clocks"])
PY
done | tee gpurun_out/r3y_e2e_repeat.txt
