#!/bin/bash
# round 2, call 3K (1 GPU): cmul as FMUL2 + FFMA2, batched twiddle loads, 32-bit store epilogue of row_inv: the whole GPU suite, c5 kernel times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r3k_tests.txt
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-pageable > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err; tail -c 300 gpurun_out/r3k_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3k_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), [(k["kernel"][:12], round(k["avg_ms"],3)) for k in d["kernels"][:3]], d["parity_spot_check"])
for s in d["other_shapes"]: print(s["shape"][:70], round(s["us_per_call"],1), "us", s.get("kernel_us_per_call"))
PY
