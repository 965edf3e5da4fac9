"""TEST INFRASTRUCTURE ONLY: compile the kernel bodies of ndarray-conv_b200/csrc as plain C++
(-DNDCONV_HOST_EMUL: one "thread" per block, blocks run in a loop) so their index logic can be checked
against the oracle on a box with no GPU.  The result answers ndconv_is_emulation() == 1 and the product
loader refuses it; nothing in the product ever builds or loads it."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "ndarray-conv_b200" / "csrc"
OUT = HERE / "_build" / "libndconv_emul.so"


def build(force: bool = False) -> Path:
    deps = list(CSRC.glob("*")) + [ROOT / "include" / "ndconv.h", Path(__file__)]
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DNDCONV_HOST_EMUL", "-x", "c++",
           str(CSRC / "api.cu"), str(CSRC / "host_logic.cpp"), "-o", str(OUT)]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
