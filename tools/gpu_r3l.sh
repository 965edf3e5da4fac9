#!/bin/bash
# round 2, call 3L (1 GPU): ncu --set full + source of the round-2e kernels on the full c5, then the slot-ring bound again (kernels now DRAM-bound)
mkdir -p gpurun_out /tmp/rep
K='regex:row_fwd|col_pass|row_inv'
ncu --set full --clock-control none --import-source on -k "$K" -s 6 -c 3 -o /tmp/rep/c5full -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e --no-pageable > gpurun_out/ncu_c5full_r02f.log 2>&1
python tools/summarize_ncu.py /tmp/rep/c5full.ncu-rep gpurun_out/r02f_ncu_full_c5 | cut -c1-600
ncu -i /tmp/rep/c5full.ncu-rep --page source --csv > gpurun_out/r02f_c5full_source.csv 2>/dev/null
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable --no-e2e > gpurun_out/r3l.json 2> gpurun_out/r3l.err || tail -3 gpurun_out/r3l.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3l.json").read().strip().splitlines()[-1])
ks=" | ".join("%s %.3f ms" % (k["kernel"], k["avg_ms"]) for k in d["kernels"] if not k["kernel"].startswith("tail"))
print(sys.argv[1], "| step %.3f ms |" % d["ms_per_step"], ks, "| spot rel %.2e" % d["parity_spot_check"]["rel"])
PY
}
{
run "product library"
cp tools/exp/bin/libndconv_cuda_ring.so ndarray-conv_b200/libndconv_cuda.so
for r in 4 8; do NDCONV_EXP_RING_TILES=$r run "ring $r tiles"; done
} | tee gpurun_out/r3l_ring_bound.txt
