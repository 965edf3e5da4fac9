#!/bin/bash
# round 2, call 3A (1 GPU): row_fwd with the inline border-pair path: parity of the fast paths, small-shape call times, c5 kernel times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_opt.py tests/test_parity_f64_fast.py tests/test_col_kres.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r3a_tests.txt
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; tail -c 300 gpurun_out/r3a_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3a_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), [(k["kernel"][:12], round(k["avg_ms"],3)) for k in d["kernels"][:3]])
for s in d["other_shapes"]: print(s["shape"][:70], round(s["us_per_call"],1), "us", s.get("kernel_us_per_call"))
PY
