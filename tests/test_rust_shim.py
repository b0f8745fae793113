"""The Rust-side binding (integration/rust) cannot be compiled here (no rustc); what can be checked is
that the generated FFI file is in step with include/ptrs_b200.h and with the ctypes mirror the tests actually use."""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generated_ffi_is_up_to_date():
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"]).returncode == 0, \
        "include/ptrs_b200.h changed: run python tools/gen_rust_ffi.py"


def test_ffi_declares_every_export_and_matches_ctypes_layouts(host):
    import pathtracer_rs_b200._abi as abi
    import pathtracer_rs_b200.gpu as gpu

    text = open(os.path.join(ROOT, "integration", "rust", "ffi.rs")).read()
    for sym in gpu.EXPORTS:
        assert re.search(rf"pub fn {sym}\(", text), f"{sym} missing from ffi.rs"
    size = {"i32": 4, "u32": 4, "f32": 4, "u64": 8, "u16": 2, "u8": 1}

    def rust_size(fields):
        # repr(C) layout with natural alignment, pointers = 8
        off, align_max = 0, 1
        for ty in fields:
            m = re.match(r"\[(\w+); (\d+)\]", ty)
            if m:
                sz, n = size[m.group(1)], int(m.group(2))
            elif ty.startswith("*"):
                sz, n = 8, 1
            else:
                sz, n = size[ty], 1
            off = (off + sz - 1) // sz * sz + sz * n
            align_max = max(align_max, sz)
        return (off + align_max - 1) // align_max * align_max

    for name in ("PtrsRay", "PtrsHit", "PtrsBvhNode", "PtrsMesh", "PtrsTexture", "PtrsMipMap", "PtrsMaterial", "PtrsLight", "PtrsEnvLight",
                 "PtrsSceneDesc", "PtrsCamera", "PtrsRenderParams", "PtrsStats"):
        body = re.search(rf"pub struct {name} \{{(.*?)\n\}}", text, flags=re.S).group(1)
        fields = re.findall(r"pub \w+: ([^,]+),", body)
        assert rust_size(fields) == C.sizeof(getattr(abi, name)), name
    shim = open(os.path.join(ROOT, "integration", "rust", "b200.rs")).read()
    for sym in re.findall(r"ffi::(ptrs_\w+)", shim):
        assert sym in gpu.EXPORTS, f"b200.rs calls {sym}, which the library does not export"


def test_shim_sources_are_self_consistent():
    """b200.rs, tables.rs and reference_additions.rs name each other's items consistently: every `Tables` method b200.rs
    calls exists in tables.rs, every exporter / accessor either file calls on a reference type is one of the additions,
    and every literal of an FFI struct lists exactly the fields ffi.rs declares."""
    rust = os.path.join(ROOT, "integration", "rust")
    b200, tables, adds, ffi = (open(os.path.join(rust, f)).read() for f in ("b200.rs", "tables.rs", "reference_additions.rs", "ffi.rs"))
    defined = set(re.findall(r"pub fn (\w+)", tables))
    for m in set(re.findall(r"\btables\.(\w+)\(", b200)) | set(re.findall(r"Tables::(\w+)\(", b200)):
        assert m in defined, f"b200.rs calls Tables::{m}, which tables.rs does not define"
    assert "push_env" in re.findall(r"pub fn (\w+)", b200)  # called back from tables.rs
    added = set(re.findall(r"fn (\w+)", adds))
    for name in ("export_flat", "offset", "get_shape", "get_material_arc", "mesh", "indices", "export", "params", "levels", "parts", "bvh"):
        assert name in added, f"reference_additions.rs lacks {name}"
        assert re.search(rf"\.{name}\(", b200 + tables), f"{name} is added to the reference but never used"
    for struct in ("PtrsEnvLight", "PtrsMaterial", "PtrsTexture", "PtrsMesh", "PtrsCamera", "PtrsBvhNode", "PtrsSceneDesc"):
        body = re.search(rf"pub struct {struct} \{{(.*?)\n\}}", ffi, flags=re.S).group(1)
        fields = set(re.findall(r"pub (\w+):", body))
        for src in (b200, tables):
            for lit in re.finditer(rf"(?<!-> )ffi::{struct} \{{(.*?)\}}[;,)\n]", src, flags=re.S):
                text = lit.group(1)
                if "zeroed" in text:
                    continue
                missing = fields - set(re.findall(r"\b[a-z_][a-z_0-9]*\b", text))
                assert not missing, f"{struct} literal lacks {sorted(missing)}"
