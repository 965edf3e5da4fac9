#!/bin/bash
# e2e (host-buffer) time of the c5 bench under environment variants
for v in "$@"; do
  env $v python bench.py --steps 6 --warmup 3 --no-cpu --no-shapes > gpurun_out/var.json 2> gpurun_out/var.err || tail -3 gpurun_out/var.err
  python - "$v" <<PY
import json,sys
d=json.loads(open("gpurun_out/var.json").read().strip().splitlines()[-1])
print(sys.argv[1], "| e2e ms/step", round(d["e2e"]["ms_per_step"],2), "Gs/s", round(d["e2e"]["value"],2))
PY
done
