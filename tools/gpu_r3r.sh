#!/bin/bash
# round 2, call 3R (2 GPUs): the multi-device paths with the round-2f library: GPU tests that need two devices, the 2-rank bench (device-timed, e2e, assembled check)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_sharded_call.py tests/test_sharded_device.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3r_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3r_n2.json 2> gpurun_out/r3r_n2.err; tail -c 300 gpurun_out/r3r_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3r_n2.json").read().strip().splitlines()[-1])
print("N=2 Gsamples/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), d.get("assembled_check"), d["config"].get("parallelism"))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
