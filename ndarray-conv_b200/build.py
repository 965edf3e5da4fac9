"""Build libndconv_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libndconv_cuda.so"
SOURCES = [CSRC / "api.cu", CSRC / "host_logic.cpp"]
HEADERS = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.inc")) + [PKG.parent / "include" / "ndconv.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=true",
              "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas", "-shared", "--expt-relaxed-constexpr"]


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    deps = SOURCES + HEADERS + [Path(__file__)]
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, *( ["-Xptxas", "-v"] if verbose else []), "-o", str(OUT), *map(str, SOURCES)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
