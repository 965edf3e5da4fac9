#!/bin/bash
# round 2, call 3H (1 GPU): upper bound of an L2-resident (fused) pipeline -- the three c5 passes with every tile's spectra in slot
# tile % ring of the workspace (experimental library built with -DNDCONV_EXP_RING; results are garbage, timings are the point)
mkdir -p gpurun_out
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable --no-e2e > gpurun_out/r3h.json 2> gpurun_out/r3h.err || tail -3 gpurun_out/r3h.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3h.json").read().strip().splitlines()[-1])
ks=" | ".join("%s %.3f ms" % (k["kernel"], k["avg_ms"]) for k in d["kernels"] if not k["kernel"].startswith("tail"))
print(sys.argv[1], "| step %.3f ms |" % d["ms_per_step"], ks, "| spot rel %.2e" % d["parity_spot_check"]["rel"])
PY
}
{
run "product library"
cp tools/exp/bin/libndconv_cuda_ring.so ndarray-conv_b200/libndconv_cuda.so
run "ring library, ring off"
for r in 2 4 8 12; do NDCONV_EXP_RING_TILES=$r run "ring $r tiles"; done
} | tee gpurun_out/r3h_ring_bound.txt
