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
