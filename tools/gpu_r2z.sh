#!/bin/bash
# round 2, call Z (1 GPU): the driver's end-of-round sequence: GPU test suite, smoke, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2z_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2z_smoke.txt
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 400 gpurun_out/r2z_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
print("Gsamples/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "rel err", d["parity_spot_check"]["rel"], "e2e", round(d["e2e"]["value"],2), "launches", d["gpu_launches"])
print([(k["kernel"][:12], round(k["avg_ms"],3), round(k["frac_of_peak"],3)) for k in d["kernels"]])
for s in d["other_shapes"]: print(s["shape"][:70], round(s["us_per_call"],1), "us", round(s["Gsamples_per_s"],1))
PY
