#!/bin/bash
# round 2, call 3W (1 GPU): pageable-array e2e leg against the number of host copy threads
mkdir -p gpurun_out
nproc
for t in 8 12 14 16; do
  NDCONV_HOST_COPY_THREADS=$t python bench.py --steps 5 --warmup 3 --no-cpu --no-shapes > gpurun_out/r3w.json 2> gpurun_out/r3w.err || tail -3 gpurun_out/r3w.err
  python - $t <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3w.json").read().strip().splitlines()[-1])
e=d["e2e"]; q=d.get("e2e_pageable") or {}
print("copy threads", sys.argv[1], "| pinned e2e %.1f ms" % e["ms_per_step"], "| pageable", {k:(round(v,2) if isinstance(v,float) else v) for k,v in q.items() if k in ("value","ms_per_step")})
PY
done | tee gpurun_out/r3w_pageable_threads.txt
