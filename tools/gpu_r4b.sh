#!/bin/bash
# round 2, call 4B (1 GPU): batched two-phase edge fetch of row_fwd for small launches: parity (small shapes, every border type), per-shape call times, c5 unchanged
# (the switch NDCONV_ROW_EDGE_BATCHED and the code path behind it were measured here and removed again; see profiles/r02f_edge_batched_ab.txt)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_opt.py tests/test_parity_small.py tests/test_baseline_configs.py tests/test_batch_fold.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r4b_tests.txt
NDCONV_ROW_EDGE_BATCHED=1 timeout 1200 python -m pytest tests/test_parity_opt.py tests/test_baseline_configs.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/r4b_tests.txt
for v in "NDCONV_ROW_EDGE_BATCHED=0" ""; do
  echo "== ${v:-default}"
  env $v python tools/run_c2_variants.py 2>&1 | cut -c1-200
  env $v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-pageable > gpurun_out/r4b.json 2> gpurun_out/r4b.err || tail -3 gpurun_out/r4b.err
  python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4b.json").read().strip().splitlines()[-1])
print("c5 step %.3f ms" % d["ms_per_step"], [(k["kernel"][:12], round(k["avg_ms"],3)) for k in d["kernels"][:3]])
for s in d["other_shapes"]: print("  ", s["shape"][:66], round(s["us_per_call"],1), "us")
PY
done | tee gpurun_out/r4b_edge_batched.txt
