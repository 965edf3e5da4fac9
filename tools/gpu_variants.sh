#!/bin/bash
# kernel times of the c5 bench under environment variants:  tools/gpu_variants.sh "A=1" "B=2 C=3" ...
for v in "$@"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-e2e > gpurun_out/var.json 2> gpurun_out/var.err || tail -3 gpurun_out/var.err
  python - "$v" <<PY
import json,sys
d=json.loads(open("gpurun_out/var.json").read().strip().splitlines()[-1])
print(sys.argv[1], "| ms/step", round(d["ms_per_step"],3), "rel err %.2e" % d["parity_spot_check"]["rel"], [(k["kernel"][:8], round(k["avg_ms"],3)) for k in d["kernels"]])
PY
done
