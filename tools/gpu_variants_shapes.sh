#!/bin/bash
# c5 step time and the per-shape call times (other_shapes) of the bench under environment variants
for v in "$@"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/var.json 2> gpurun_out/var.err || tail -3 gpurun_out/var.err
  python - "$v" <<PY
import json,sys
d=json.loads(open("gpurun_out/var.json").read().strip().splitlines()[-1])
print(sys.argv[1], "| c5 ms/step", round(d["ms_per_step"],3), "rel err %.2e" % d["parity_spot_check"]["rel"], "| us/call", [(s["shape"][:8], round(s["us_per_call"],2)) for s in d["other_shapes"]])
PY
done
