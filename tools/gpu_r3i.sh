#!/bin/bash
# round 2, call 3I (1 GPU): per-kernel times of the slab one rank of an N-rank run gets (bench.py --as-slab W:R)
mkdir -p gpurun_out
for sl in 1:0 2:0 4:1 8:0 8:3 8:7; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable --no-e2e --no-verify --as-slab $sl > gpurun_out/r3i.json 2> gpurun_out/r3i.err || tail -3 gpurun_out/r3i.err
  python - $sl <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3i.json").read().strip().splitlines()[-1])
ks=" | ".join("%s %.3f" % (k["kernel"].replace("row_fwd_pad_r2c","fwd").replace("col_fwd_mul_inv","col").replace("row_inv_c2r_crop","inv"), k["avg_ms"]) for k in d["kernels"])
print("slab", sys.argv[1], "| step %.3f ms |" % d["ms_per_step"], ks, "| spot rel %.2e" % d["parity_spot_check"]["rel"])
PY
done | tee gpurun_out/r3i_slab_kernels.txt
