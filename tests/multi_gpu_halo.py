"""torchrun script (one rank per GPU, NCCL): device-resident, row-partitioned conv_fft with NCCL halo exchange must equal
the single-GPU result computed on rank 0.  Launched by tests/test_multi_gpu.py when >= 2 GPUs are visible."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("ndarray-conv_b200")
    sharded = importlib.import_module("ndarray-conv_b200.sharded")
    proc = pkg.get_fft_processor(local)
    ok = True
    for padding, shape, ks in ((pkg.PaddingMode.Reflect, (3000, 2600), (31, 17)), (pkg.PaddingMode.Circular, (1500, 300), (9, 5)),
                               (pkg.PaddingMode.Const(2.0), (2500, 2100), (5, 63))):
        x = np.random.default_rng(9).random(shape, dtype=np.float32)
        k = np.random.default_rng(10).random(ks, dtype=np.float32)
        b = sharded.row_partition(shape[0], world)
        x_local = torch.from_numpy(x[b[rank]:b[rank + 1]].copy()).to(dev)
        y, ob, oe = sharded.conv_fft_device_resident(x_local, shape[0], k, pkg.ConvMode.Full, padding, rank, world, proc)
        torch.cuda.synchronize(dev)
        ref = pkg.conv_fft_with_processor(x, k, pkg.ConvMode.Full, padding, proc)[ob:oe]     # single-GPU result, same rows
        err = float(np.max(np.abs(y.cpu().numpy() - ref)))
        tol = 8 * np.finfo(np.float32).eps * np.log2(1024 * 2048) * float(np.max(np.abs(ref)))
        ok = ok and (y.shape[0] == oe - ob) and err <= tol
        if rank == 0:
            print(f"halo-exchange conv_fft {shape} k={ks}: rows [{ob},{oe}) max|diff|={err:.3e} tol={tol:.3e}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    proc.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_HALO_OK" if int(flag.item()) == 1 else "MULTI_GPU_HALO_FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
