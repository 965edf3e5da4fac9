#!/bin/bash
# round 2, call 3S (8 GPUs): the sharded bench at N = 8 and N = 4 with the round-2f library (device-timed, e2e, assembled check)
mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r3s_n$n.json 2> gpurun_out/r3s_n$n.err; tail -c 200 gpurun_out/r3s_n$n.err
python - $n <<PY
import json,sys
n=sys.argv[1]
d=json.loads(open("gpurun_out/r3s_n%s.json" % n).read().strip().splitlines()[-1])
print("N=%s Gsamples/s" % n, round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "e2e ms", round(d["e2e"]["ms_per_step"],1), "copy floor ms", d["e2e"].get("copy_floor_ms"), (d.get("assembled_check") or {}).get("ok"), d["clocks"])
PY
done | tee gpurun_out/r3s_scale.txt
