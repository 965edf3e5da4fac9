#!/bin/bash
# round 2, call 3G (1 GPU): e2e (host buffers) against the slab size of the pipelined host path
mkdir -p gpurun_out
for mb in 32 64 128 256; do
  NDCONV_PIPE_SLAB_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu --no-shapes --no-pageable > gpurun_out/r3g.json 2> gpurun_out/r3g.err || tail -3 gpurun_out/r3g.err
  python - $mb <<PY
import json,sys
d=json.loads(open("gpurun_out/r3g.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("slab MB", sys.argv[1], "e2e ms", round(e["ms_per_step"],2), "Gsamples/s", round(e["value"],2), "copy floor ms", round(e.get("copy_floor_ms",0),2), "frac", round(e.get("frac_of_copy_floor",0),3))
PY
done | tee gpurun_out/r3g_e2e_slab.txt
