#!/bin/bash
# round 2, call 3M (1 GPU): Tensor-Memory column pass with register -> global stores (TMA moves the loads only) against TMA stores
mkdir -p gpurun_out
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable --no-e2e > gpurun_out/r3m.json 2> gpurun_out/r3m.err || tail -3 gpurun_out/r3m.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3m.json").read().strip().splitlines()[-1])
ks=" | ".join("%s %.3f ms" % (k["kernel"], k["avg_ms"]) for k in d["kernels"] if not k["kernel"].startswith("tail"))
print(sys.argv[1], "| step %.3f ms |" % d["ms_per_step"], ks, "| spot rel %.2e" % d["parity_spot_check"]["rel"])
PY
}
{
run "TMA stores (default)"
NDCONV_COL_STG=1 run "NDCONV_COL_STG=1"
run "TMA stores (default)"
NDCONV_COL_STG=1 run "NDCONV_COL_STG=1"
NDCONV_COL_STG=1 timeout 600 python -m pytest tests/test_col_kres.py tests/test_baseline_configs.py -m gpu -x -q 2>&1 | tail -3
} | tee gpurun_out/r3m_col_stg.txt
# c2: where do row_fwd's 19 us go?  (ncu --set full + source, one launch of each kernel)
mkdir -p /tmp/rep
ncu --set full --clock-control none --import-source on -k "regex:row_fwd|col_pass|row_inv" -s 9 -c 3 -o /tmp/rep/c2 -f python tools/run_c2.py > gpurun_out/ncu_c2_r02f.log 2>&1
tail -2 gpurun_out/ncu_c2_r02f.log
python tools/summarize_ncu.py /tmp/rep/c2.ncu-rep gpurun_out/r02f_ncu_full_c2 | cut -c1-500
ncu -i /tmp/rep/c2.ncu-rep --page source --csv > gpurun_out/r02f_c2_source.csv 2>/dev/null
