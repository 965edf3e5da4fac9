#!/bin/bash
# round 2, call F (1 GPU): batch fold tests + the whole GPU suite + a full bench line with the new other_shapes
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "not full_bands" ) 2>&1 | tail -6
( time python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2f_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
print("Gsamples/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],2))
for s in d["other_shapes"]:
    print(s["shape"][:60], "| us", round(s["us_per_call"],1), "| per problem", round(s.get("us_per_problem",0),2), "| GMAC/s", round(s.get("GMAC_per_s",0),1), "| GB/s", round(s["compulsory_GBps"],1), s["kernel_us_per_call"])
PY
