"""-m gpu, needs >= 2 B200: NCCL halo exchange for device-resident, row-partitioned input."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_device_resident_halo_exchange_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (covered on CPU by tests/test_sharded_gloo.py with gloo)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", str(ROOT / "tests" / "multi_gpu_halo.py")], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert "MULTI_GPU_HALO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
