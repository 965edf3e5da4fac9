#!/bin/bash
# round 2, call 3T (1 GPU): f64 fast path with batched twiddle loads (twiddle_store_d): parity, per-kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_f64_fast.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r3t_tests.txt
python tools/run_f64_shapes.py 2>&1 | cut -c1-420 | tee gpurun_out/r3t_f64.txt
# row_fwd with multiply + shift row resolution (FastDiv): fast-path parity, c5 kernel times
timeout 1200 python -m pytest tests/test_parity_opt.py tests/test_baseline_configs.py tests/test_batch_fold.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/r3t_tests.txt
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable --no-e2e > gpurun_out/r3t.json 2> gpurun_out/r3t.err || tail -3 gpurun_out/r3t.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3t.json").read().strip().splitlines()[-1])
print("c5 step %.3f ms |" % d["ms_per_step"], " | ".join("%s %.3f ms" % (k["kernel"], k["avg_ms"]) for k in d["kernels"] if not k["kernel"].startswith("tail")), "| spot rel %.2e" % d["parity_spot_check"]["rel"])
PY
done | tee gpurun_out/r3t_c5.txt
