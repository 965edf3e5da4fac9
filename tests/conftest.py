import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def pkg():
    """the product's Python host mirror (directory name has a hyphen -> importlib)"""
    import importlib
    return importlib.import_module("ndarray-conv_b200")


@pytest.fixture(scope="session")
def emul_lib(pkg):
    """TEST-ONLY host emulation of the kernel bodies (tests/emul/build_emul.py)."""
    import ctypes
    sys.path.insert(0, str(ROOT / "tests" / "emul"))
    import build_emul
    return pkg.Library(ctypes.CDLL(str(build_emul.build())))


@pytest.fixture(scope="session")
def cuda_lib(pkg):
    return pkg.get_library()


@pytest.fixture(params=["emul", pytest.param("cuda", marks=pytest.mark.gpu)])
def ndc(request, pkg):
    """(package, library): the kernel bodies under host emulation on a CPU box, the real sm_100a build on a B200."""
    lib = request.getfixturevalue("emul_lib" if request.param == "emul" else "cuda_lib")
    return pkg, lib
