"""g++ the C++ host-mirror tests against the in-tree libndconv_cuda.so (rpath = its directory)."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIBDIR = ROOT / "ndarray-conv_b200"
OUT = HERE / "_build" / "test_ndconv_hpp"


def build(force=False):
    src = HERE / "test_ndconv_hpp.cpp"
    deps = [src, ROOT / "include" / "ndconv.hpp", ROOT / "include" / "ndconv.h", LIBDIR / "libndconv_cuda.so"]
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", str(src), "-o", str(OUT), f"-L{LIBDIR}", "-lndconv_cuda", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
