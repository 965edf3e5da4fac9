#!/bin/bash
# one gpurun call: fast-path parity tests + c5 bench line (kernel times) -> gpurun_out/<tag>.json
tag=${1:-check}
python -m pytest tests/test_parity_opt.py tests/test_baseline_configs.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("Gsamples/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "rel err", d["parity_spot_check"]["rel"])
print([(k["kernel"][:8], round(k["avg_ms"],3), round(k["frac_of_peak"],3)) for k in d["kernels"]])
print("e2e", round(d["e2e"]["value"],2), d["clocks"])
PY
