#!/bin/bash
# kernel times of several bench workloads:  tools/gpu_workloads.sh c5 s512 ...
for v in "$@"; do
  python bench.py --workload $v --steps 50 --warmup 5 --no-cpu --no-shapes --no-e2e > gpurun_out/var.json 2> gpurun_out/var.err || tail -3 gpurun_out/var.err
  python - "$v" <<PY
import json,sys
d=json.loads(open("gpurun_out/var.json").read().strip().splitlines()[-1])
print(sys.argv[1], "| ms/step", round(d["ms_per_step"],4), "Gs/s", round(d["value"],1), "rel err %.2e" % d["parity_spot_check"]["rel"], [(k["kernel"][:8], round(k["avg_ms"],4)) for k in d["kernels"]], "ws MB", d["workspace_bytes"]/1e6)
PY
done
