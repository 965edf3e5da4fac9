#!/bin/bash
# round 2, call 4D (1 GPU): zero rows that share a warp with sample rows (T < 32) skip the per-column loop: parity, c2-shaped calls by border type, small shapes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_opt.py tests/test_parity_small.py tests/test_baseline_configs.py tests/test_batch_fold.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r4d_tests.txt
{
python tools/run_c2_variants.py 2>&1 | cut -c1-200
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-pageable > gpurun_out/r4d.json 2> gpurun_out/r4d.err || tail -3 gpurun_out/r4d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4d.json").read().strip().splitlines()[-1])
print("c5 step %.3f ms" % d["ms_per_step"], [(k["kernel"][:12], round(k["avg_ms"],3)) for k in d["kernels"][:3]])
for s in d["other_shapes"]: print("  ", s["shape"][:66], round(s["us_per_call"],1), "us")
PY
} | tee gpurun_out/r4d_zero_rows.txt
