#!/bin/bash
# round 2, call 3N (1 GPU): what bounds col_pass_tma_kres?  experimental library (-DNDCONV_EXP_RING): workspace loads off / stores off / both off
mkdir -p gpurun_out
# c4 on the register-blocked direct variant (size gate lowered) against the tap-list kernel
for v in "" "NDCONV_BLOCKED_MIN_OUT=1"; do
  env $v python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-pageable > gpurun_out/r3n_c4.json 2> gpurun_out/r3n_c4.err || tail -3 gpurun_out/r3n_c4.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3n_c4.json").read().strip().splitlines()[-1])
for s in d["other_shapes"]:
    if s["shape"].startswith("c4"): print("c4", sys.argv[1] or "default", round(s["us_per_call"],1), "us", s.get("kernel_us_per_call"))
PY
done | tee gpurun_out/r3n_c4.txt
cp tools/exp/bin/libndconv_cuda_ring.so ndarray-conv_b200/libndconv_cuda.so
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable --no-e2e > gpurun_out/r3n.json 2> gpurun_out/r3n.err || tail -3 gpurun_out/r3n.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3n.json").read().strip().splitlines()[-1])
ks=" | ".join("%s %.3f ms" % (k["kernel"], k["avg_ms"]) for k in d["kernels"] if not k["kernel"].startswith("tail"))
print(sys.argv[1], "| step %.3f ms |" % d["ms_per_step"], ks)
PY
}
{
run "all on"
NDCONV_EXP_FLAGS=1 run "no workspace loads"
NDCONV_EXP_FLAGS=2 run "no stores"
NDCONV_EXP_FLAGS=3 run "neither (compute + kernel-spectrum stream only)"
NDCONV_EXP_RING_TILES=4 run "ring 4"
NDCONV_EXP_RING_TILES=4 NDCONV_EXP_FLAGS=2 run "ring 4, no stores"
} | tee gpurun_out/r3n_col_bounds.txt
