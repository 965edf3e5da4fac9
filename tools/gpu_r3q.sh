#!/bin/bash
# round 2, call 3Q (1 GPU): row_fwd tile visit order (axis-0 tile index fastest: halo rows re-read out of L2) and streaming workspace stores
# (the two switches were measured and removed again: NDCONV_ROW_ORDER / NDCONV_ROW_WS_CS no longer exist in the library; see profiles/r02f_row_order_ab.txt)
mkdir -p gpurun_out
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable --no-e2e > gpurun_out/r3q.json 2> gpurun_out/r3q.err || tail -3 gpurun_out/r3q.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r3q.json").read().strip().splitlines()[-1])
ks=" | ".join("%s %.3f ms" % (k["kernel"], k["avg_ms"]) for k in d["kernels"] if not k["kernel"].startswith("tail"))
print(sys.argv[1], "| step %.3f ms |" % d["ms_per_step"], ks, "| spot rel %.2e" % d["parity_spot_check"]["rel"])
PY
}
{
run "default"
NDCONV_ROW_ORDER=1 run "ROW_ORDER=1"
NDCONV_ROW_WS_CS=1 run "ROW_WS_CS=1"
NDCONV_ROW_ORDER=1 NDCONV_ROW_WS_CS=1 run "ROW_ORDER=1 ROW_WS_CS=1"
run "default"
for v in "" "NDCONV_ROW_ORDER=1 NDCONV_ROW_WS_CS=1"; do
env $v ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:row_fwd" -s 2 -c 1 --csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-shapes --no-e2e --no-pageable 2>/dev/null | grep -E "row_fwd" | awk -F'","' '{print $13, $15}' | tr '\n' ' '; echo " <- $v"
done
} | tee gpurun_out/r3q_row_order.txt
