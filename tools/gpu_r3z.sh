#!/bin/bash
# round 2, call 3Z (1 GPU): the default bench line of the final library (HEAD of round 2), twice
mkdir -p gpurun_out
for i in 1 2; do
  python bench.py > gpurun_out/r3z_bench_$i.json 2> gpurun_out/r3z_bench_$i.err; tail -c 200 gpurun_out/r3z_bench_$i.err
  python - $i <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3z_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
e=d["e2e"]; q=d.get("e2e_pageable") or {}
print("run", sys.argv[1], "| step %.3f ms = %.1f Gs/s" % (d["ms_per_step"], d["value"]), "| e2e %.1f ms = %.2f Gs/s" % (e["ms_per_step"], e["value"]), "floor %.1f ms" % (e.get("copy_floor_ms") or 0), "| pageable %.1f ms" % q.get("ms_per_step", 0), "| roofline", round(d["roofline"]["frac"],3), "| cpu", round(d["cpu_baseline"]["value"],3))
PY
done | tee gpurun_out/r3z_final.txt
