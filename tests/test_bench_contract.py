"""bench.py contract on a GPU-less box: the reference arm (`--impl reference`: the oracle's pocketfft restatement of the reference
pipeline on the host cores, a bounded sample of the c5 workload) prints ONE JSON line with the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gsamples/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("conv_fft output Gsamples/s") and "32768" in d["config"]["workload"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_default_arm_fails_loudly_without_a_gpu():
    """no CPU fallback: without a device our arm must error out, not print a number"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{") and '"value"' in l]
