#!/bin/bash
# round 2, call A: full GPU test suite, NDCONV_COL_EARLY A/B (parity + timing), repaired reference arm, full bench line
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
( time python -m pytest tests -m gpu -x -q -s -k "not full_bands" ) 2>&1 | tail -6
( time python -m pytest tests/test_baseline_configs.py -m gpu -x -q -s -k "full_bands" ) 2>&1 | tail -8
echo "== COL_EARLY parity"
NDCONV_COL_EARLY=1 python -m pytest tests/test_parity_opt.py tests/test_baseline_configs.py -m gpu -x -q -k "not full_bands" 2>&1 | tail -3
for v in "" "NDCONV_COL_EARLY=1"; do
  echo "== variant [$v]"
  env $v python bench.py --steps 20 --warmup 5 --no-cpu --no-shapes --no-e2e 2> gpurun_out/r2a_var.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), [(k['kernel'][:12], round(k['avg_ms'],3)) for k in d['kernels']], d['parity_spot_check']['rel'])"
done
echo "== reference arm"
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err ) 2>&1 | tail -3
cut -c1-700 gpurun_out/r2a_ref.json
echo "== bench"
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err ) 2>&1 | tail -3
python - <<PY
import json
d=json.loads(open("gpurun_out/r2a_bench.json").read().strip().splitlines()[-1])
print("Gsamples/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", d["e2e"], d["assembled_check"], d["cpu_baseline"])
PY
