#!/bin/bash
# round 2, call 4E (1 GPU): existing switches re-checked on the round-2f kernels (row_fwd L2 prefetch position, twiddles of the column pass from shared memory)
mkdir -p gpurun_out
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable --no-e2e > gpurun_out/r4e.json 2> gpurun_out/r4e.err || tail -3 gpurun_out/r4e.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r4e.json").read().strip().splitlines()[-1])
ks=" | ".join("%s %.3f ms" % (k["kernel"], k["avg_ms"]) for k in d["kernels"] if not k["kernel"].startswith("tail"))
print(sys.argv[1], "| step %.3f ms |" % d["ms_per_step"], ks)
PY
}
{
run "default (NDCONV_ROW_PF=2)"
NDCONV_ROW_PF=0 run "NDCONV_ROW_PF=0"
NDCONV_ROW_PF=1 run "NDCONV_ROW_PF=1"
NDCONV_COL_KRES_NO_TWT=1 run "NDCONV_COL_KRES_NO_TWT=1"
NDCONV_DISABLE_PDL=1 run "NDCONV_DISABLE_PDL=1"
run "default"
} | tee gpurun_out/r4e_switches.txt
