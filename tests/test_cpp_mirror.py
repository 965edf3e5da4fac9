"""include/ndconv.hpp (C++ mirror of the crate's trait API): compiled tests that read like the reference's own."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "cpp"))


def _binary():
    import importlib
    importlib.import_module("ndarray-conv_b200.build").build_cuda()
    import build_cpp_tests
    return build_cpp_tests.build()


def test_cpp_host_only_checks():
    r = subprocess.run([str(_binary()), "--host-only"], capture_output=True, text=True, timeout=120)
    assert "HOST_CPP_TESTS_PASSED" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_reference_style_tests_on_device():
    r = subprocess.run([str(_binary())], capture_output=True, text=True, timeout=300)
    assert "ALL_CPP_TESTS_PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
