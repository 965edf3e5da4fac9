"""The C-ABI library loads on a CPU box and exports every symbol include/ndconv.h declares (no compute calls)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported(pkg):
    import importlib
    importlib.import_module("ndarray-conv_b200.build").build_cuda()
    lib = pkg.get_library()
    header = (ROOT / "include" / "ndconv.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)      # strip comments
    declared = set(re.findall(r"\b(ndconv_[a-z_0-9]+)\s*\(", header))
    assert declared == set(pkg.EXPORTED_SYMBOLS), declared ^ set(pkg.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib.c, sym), sym
    assert not lib.is_emulation
    assert b"sm_100a" in lib.c.ndconv_version()


def test_host_logic_without_gpu(pkg):
    lib = pkg.get_library()
    assert lib.c.ndconv_good_fft_size(5030) == 5120 and lib.c.ndconv_good_fft_size(32892) == 36864
    assert lib.c.ndconv_plan_fft_size(32892, 1) == 32928
    pads, strides = pkg.ConvMode.Same.unfold([4, 3], [1, 2], lib)
    assert pads.tolist() == [[2, 1], [2, 2]] and strides.tolist() == [1, 1]
    assert lib.c.ndconv_dtype_size(5) == 16


def test_no_cpu_fallback(pkg):
    """Without a device the compute entry points must fail loudly (NDCONV_ERR_CUDA), never compute on the host."""
    import numpy as np
    lib = pkg.get_library()
    if lib.c.ndconv_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.NdConvError) as e:
        pkg.conv(np.ones(4, np.int32), np.ones(2, np.int32))
    assert e.value.status == pkg.ERR_CUDA
    with pytest.raises(pkg.NdConvError) as e:
        pkg.conv_fft(np.ones(4, np.float32), np.ones(2, np.float32))
    assert e.value.status == pkg.ERR_CUDA
    # shape errors still surface first, exactly as in the reference
    with pytest.raises(pkg.NdConvError) as e:
        pkg.conv(np.ones(3, np.int32), np.ones(5, np.int32), pkg.ConvMode.Valid)
    assert e.value.status == pkg.ERR_MISMATCH_SHAPE


def test_product_loader_rejects_emulation(pkg, emul_lib):
    assert emul_lib.is_emulation
    src = (ROOT / "ndarray-conv_b200" / "__init__.py").read_text()
    assert "refusing to load a host-emulation build" in src
    assert "oracle" not in "".join(l for l in src.splitlines() if l.strip().startswith(("import", "from")))


def test_product_library_is_not_the_experimental_build(pkg):
    """tools/exp/build_ring_lib.sh builds a library with -DNDCONV_EXP_RING whose results are wrong by construction (timing experiments
    of DESIGN.md section 9).  The in-tree product library must never be that build: its environment switches do not exist in it."""
    blob = Path(pkg.LIB_PATH).read_bytes()
    assert b"NDCONV_EXP_RING_TILES" not in blob and b"NDCONV_EXP_FLAGS" not in blob
    assert b"ndconv_conv_fft" in blob
