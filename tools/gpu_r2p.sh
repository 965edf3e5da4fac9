#!/bin/bash
# round 2, call P (1 GPU): full bench line + ncu evidence (launch list, DRAM traffic, --set full) of the TMEM-resident column pass and the new row_fwd
mkdir -p gpurun_out
python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -c 600 gpurun_out/r2p_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
print("Gsamples/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "rel err", d["parity_spot_check"]["rel"], "e2e", round(d["e2e"]["value"],2))
print([(k["kernel"][:12], round(k["avg_ms"],3), round(k["frac_of_peak"],3)) for k in d["kernels"]])
print(d["roofline"]); print(d["clocks"])
PY
bash tools/gpu_profile.sh r02b
