#!/bin/bash
# round 2, call G (8 GPUs): host<->device ceiling of the box (which GPUs share an uplink), e2e of the bench at N = 8 / 4 / 2 on
# linear and spread GPU picks, the device-resident sharded call at 2 / 4 / 8 GPUs, the multi-GPU tests
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m 2>/dev/null | sed 's/\x1b\[[0-9;]*m//g' | head -14 > gpurun_out/r2g_topo.txt; nproc >> gpurun_out/r2g_topo.txt; free -g | head -2 >> gpurun_out/r2g_topo.txt
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r2g_topo.txt
echo "== probe"
( time $TR --nproc-per-node 8 --master-port 29511 tools/pcie_probe_multi.py > gpurun_out/r2g_pcie_probe.jsonl 2> gpurun_out/r2g_probe.err ) 2>&1 | grep real
cat gpurun_out/r2g_pcie_probe.jsonl | cut -c1-330
summ() { python - "$1" "$2" <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "| N", d["n_gpus"], "| ms/step", round(d["ms_per_step"],3), "| Gs/s", round(d["value"],1), "| e2e ms", round(d["e2e"]["ms_per_step"],1), "Gs/s", round(d["e2e"]["value"],2), "|", d.get("assembled_check"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
B="bench.py --steps 10 --warmup 3 --no-cpu --no-shapes --no-pageable"
echo "== bench e2e"
( time $TR --nproc-per-node 8 --master-port 29512 $B --gpus 8 > gpurun_out/r2g_n8.json 2> gpurun_out/r2g_n8.err ) 2>&1 | grep real; summ "N=8" gpurun_out/r2g_n8.json
NDCONV_PIPE_SLAB_MB=16 $TR --nproc-per-node 8 --master-port 29513 $B --gpus 8 --no-verify > gpurun_out/r2g_n8_s16.json 2> gpurun_out/r2g_n8_s16.err; summ "N=8 slab16" gpurun_out/r2g_n8_s16.json
$TR --nproc-per-node 4 --master-port 29514 $B --gpus 4 --no-verify > gpurun_out/r2g_n4.json 2> gpurun_out/r2g_n4.err; summ "N=4 gpus 0-3" gpurun_out/r2g_n4.json
CUDA_VISIBLE_DEVICES=0,2,4,6 $TR --nproc-per-node 4 --master-port 29515 $B --gpus 4 --no-verify > gpurun_out/r2g_n4s.json 2> gpurun_out/r2g_n4s.err; summ "N=4 gpus 0,2,4,6" gpurun_out/r2g_n4s.json
$TR --nproc-per-node 2 --master-port 29516 $B --gpus 2 --no-verify > gpurun_out/r2g_n2.json 2> gpurun_out/r2g_n2.err; summ "N=2 gpus 0,1" gpurun_out/r2g_n2.json
CUDA_VISIBLE_DEVICES=0,4 $TR --nproc-per-node 2 --master-port 29517 $B --gpus 2 --no-verify > gpurun_out/r2g_n2s.json 2> gpurun_out/r2g_n2s.err; summ "N=2 gpus 0,4" gpurun_out/r2g_n2s.json
echo "== device-resident sharded call"
for n in 2 4 8; do timeout 300 python tools/run_sharded_device.py 32768 $n 2>&1 | tail -1 | tee -a gpurun_out/r2g_sharded_device.jsonl | cut -c1-420; done
echo "== multi-GPU tests"
( time timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_sharded_device.py tests/test_sharded_call.py -m gpu -x -q -s ) 2>&1 | grep -E "visible GPUs|passed|failed|real" | sort | uniq -c | tail -6
