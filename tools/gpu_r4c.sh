#!/bin/bash
# round 2, call 4C (1 GPU): the in-tree library at HEAD once more: whole GPU suite, smoke, short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r4c_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r4c_smoke.txt
python bench.py --no-shapes --no-pageable > gpurun_out/r4c_bench.json 2> gpurun_out/r4c_bench.err; tail -c 200 gpurun_out/r4c_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4c_bench.json").read().strip().splitlines()[-1])
print("step %.3f ms = %.1f Gs/s | e2e %.1f ms = %.2f Gs/s (floor %.1f) | roofline %.3f | launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["copy_floor_ms"], d["roofline"]["frac"], d["gpu_launches"]))
PY
