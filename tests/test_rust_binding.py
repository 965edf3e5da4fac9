"""bindings/rust: the `cuda` feature of the crate, shipped as source (no Rust toolchain in this image).
A compiler would catch drift between include/ndconv.h and src/cuda/ffi.rs; here a parser does: every C declaration has exactly
one Rust `extern` item, `#[repr(C)]` structs have the header's field order and the ctypes mirror's offsets / sizes, enum codes
agree, and src/cuda/mod.rs only uses FFI items that exist."""
import ctypes
import importlib
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HDR = (ROOT / "include" / "ndconv.h").read_text()
FFI = (ROOT / "bindings" / "rust" / "src" / "cuda" / "ffi.rs").read_text()
MOD = (ROOT / "bindings" / "rust" / "src" / "cuda" / "mod.rs").read_text()


def strip_comments(src, rust=False):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def c_functions():
    src = strip_comments(HDR)
    return set(re.findall(r"\b(ndconv_[a-z0-9_]+)\s*\(", src))


def rust_functions():
    return set(re.findall(r"pub fn (ndconv_[a-z0-9_]+)\s*\(", strip_comments(FFI)))


def test_every_c_declaration_has_one_rust_extern():
    c, r = c_functions(), rust_functions()
    assert c == r, {"only in ndconv.h": sorted(c - r), "only in ffi.rs": sorted(r - c)}
    pkg = importlib.import_module("ndarray-conv_b200")
    assert set(pkg.EXPORTED_SYMBOLS) == c, sorted(set(pkg.EXPORTED_SYMBOLS) ^ c)


def c_struct_fields(name):
    src = strip_comments(HDR)
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S)
    assert m, name
    out = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)\s*((?:\[[^\]]+\])*)$", decl)
        ctype, names_part = decl, None
        # several declarators on one line ("int64_t out_begin, out_end")
        first = re.match(r"((?:const\s+)?(?:unsigned\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\**)\s*(.*)$", decl)
        base, rest = first.group(1).strip(), first.group(2)
        for d in rest.split(","):
            d = d.strip()
            ptr = d.startswith("*")
            d = d.lstrip("* ")
            nm = re.match(r"([A-Za-z_][A-Za-z0-9_]*)", d).group(1)
            dims = [x for x in re.findall(r"\[([^\]]+)\]", d)]
            out.append((nm, base + ("*" if ptr else ""), dims))
    return out


RUST_SIZES = {"i32": 4, "c_int": 4, "i64": 8, "u8": 1, "f64": 8, "usize": 8, "c_char": 1}


def rust_struct_fields(name):
    src = strip_comments(FFI)
    m = re.search(r"pub struct %s \{(.*?)\n\}" % name, src, flags=re.S)
    assert m, name
    out = []
    for nm, ty in re.findall(r"pub (r#\w+|\w+)\s*:\s*([^,\n]+(?:\[[^\n]*\])?)\s*,", m.group(1)):
        out.append((nm.replace("r#", ""), ty.strip()))
    return out


def rust_type_layout(ty):
    """(size, align) of a Rust FFI type under repr(C) on a 64-bit target"""
    ty = ty.strip()
    if ty.startswith("*"):
        return 8, 8
    m = re.match(r"\[(.*);\s*([A-Z_0-9a-z]+)\]$", ty)
    if m:
        n = {"MAX_DIM": 6}.get(m.group(2)) or int(m.group(2))
        s, a = rust_type_layout(m.group(1))
        return s * n, a
    if ty in RUST_SIZES:
        return RUST_SIZES[ty], RUST_SIZES[ty]
    if ty.startswith("ndconv_"):
        return rust_struct_layout(ty)[0:2]
    raise AssertionError("unknown Rust type " + ty)


def rust_struct_layout(name):
    off, align, offs = 0, 1, {}
    for nm, ty in rust_struct_fields(name):
        s, a = rust_type_layout(ty)
        off = (off + a - 1) // a * a
        offs[nm] = (off, s)
        off += s
        align = max(align, a)
    return (off + align - 1) // align * align, align, offs


@pytest.mark.parametrize("cname,ct", [("ndconv_border", "_Border"), ("ndconv_problem", "_Problem"), ("ndconv_plan_info", "_PlanInfo"), ("ndconv_slab", "_Slab"),
                                      ("ndconv_shard", "_Shard"), ("ndconv_shard_info", "_ShardInfo")])
def test_struct_layouts_agree(cname, ct):
    pkg = importlib.import_module("ndarray-conv_b200")
    cstruct = getattr(pkg, ct)
    c_fields = [f[0] for f in c_struct_fields(cname)]
    r_fields = [f[0] for f in rust_struct_fields(cname)]
    py_fields = [f[0] for f in cstruct._fields_]
    assert c_fields == r_fields == py_fields
    size, _, offs = rust_struct_layout(cname)
    assert size == ctypes.sizeof(cstruct)
    for nm in py_fields:
        d = getattr(cstruct, nm)
        assert offs[nm] == (d.offset, d.size), (nm, offs[nm], d.offset, d.size)


def test_enum_codes_agree():
    src = strip_comments(HDR)
    c_vals = {k: int(v) for k, v in re.findall(r"\bNDCONV_([A-Z0-9_]+)\s*=\s*(\d+)", src)}
    r_vals = {k: int(v) for k, v in re.findall(r"pub const ([A-Z0-9_]+)\s*:\s*\w+\s*=\s*(\d+);", strip_comments(FFI))}
    r_vals.pop("MAX_DIM")
    assert r_vals and all(k in c_vals for k in r_vals), sorted(set(r_vals) - set(c_vals))
    assert {k: c_vals[k] for k in r_vals} == r_vals
    # every status / dtype / border / memory / path code of the header is named on the Rust side
    for k in c_vals:
        if k.startswith(("MODE_", "MAX_")):
            continue                     # ConvMode is lowered on the Rust side (ConvMode::unfold), never sent across
        assert k in r_vals, k
    pkg = importlib.import_module("ndarray-conv_b200")
    assert int(re.search(r"#define NDCONV_MAX_DIM (\d+)", HDR).group(1)) == pkg.MAX_DIM == 6


def test_mod_rs_uses_only_declared_ffi_items():
    decl = rust_functions() | set(re.findall(r"pub const ([A-Z0-9_]+)", FFI)) | set(re.findall(r"pub struct (\w+)", FFI))
    used = set(re.findall(r"(?<![A-Za-z0-9_:])ffi::([A-Za-z0-9_]+)", strip_comments(MOD)))      # not std::ffi::CStr
    assert used <= decl, sorted(used - decl)
    # the trait surface of src/lib.rs:74-78 is implemented, with one signature per helper
    for needle in ("impl<'a, T, S, SK, const N: usize> ConvExt<'a, T, S, SK, N> for ArrayBase",
                   "impl<'a, T, InElem, S, SK, const N: usize> ConvFFTExt<'a, T, InElem, S, SK, N> for ArrayBase",
                   "impl<T, InElem> Processor<T, InElem> for CudaProcessor<T, InElem>", "fn conv_fft_with_processor(", "fn conv_fft_par(",
                   "fn forward<", "fn backward<", "ndconv_fft_forward", "ndconv_fft_backward", "ndconv_host_register"):
        assert needle in MOD, needle
    assert len(re.findall(r"\nfn check<", MOD)) == 1 and len(re.findall(r"\nfn lower<", MOD)) == 1
