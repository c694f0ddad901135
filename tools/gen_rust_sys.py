#!/usr/bin/env python
"""Write portello-b200-sys/src/lib.rs from include/portello_b200.h (tools/ffi_parse.py does the parsing).

The Rust toolchain is absent from this image, so the crate ships as source; tests/test_ffi_binding.py re-parses BOTH
files on every CPU test run and fails when a struct field, its order, its width or a function signature differs, so
the binding cannot rot the way a snippet in a document does.

usage: python tools/gen_rust_sys.py            (rewrites the file in place)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ffi_parse  # noqa: E402

HEADER = os.path.join(ROOT, "include", "portello_b200.h")
OUT = os.path.join(ROOT, "portello-b200-sys", "src", "lib.rs")

PRELUDE = '''//! Raw FFI declarations of `libportello_b200.so`: the B200-native read-mapping transfer ("liftover") path of portello.
//!
//! GENERATED from `include/portello_b200.h` by `tools/gen_rust_sys.py`; do not edit by hand.  The header is the contract
//! and documents every item (with the portello source lines each entry point replaces); `tests/test_ffi_binding.py`
//! checks this file against it field for field.
//!
//! Safe wrappers (`Context`, `Batch`, `LiftResult`) belong in a `portello-b200` crate on top of this one; portello's
//! `read_alignment_scanner` would call `ptl_lift_submit` / `ptl_lift_wait` from its per-window workers (INTEGRATION.md).
#![allow(non_camel_case_types)]
#![allow(clippy::too_many_arguments)]
use std::os::raw::{c_char, c_int, c_void};

'''


def main():
    m = ffi_parse.parse_header(HEADER)
    out = [PRELUDE]
    out.append("// ---------------------------------------------------------------- constants (status codes, stage masks, flags)\n")
    for k, v in m["consts"].items():
        ty = "i32" if v < 0 or k.startswith(("PTL_OK", "PTL_ERR", "PTL_REC", "PTL_PAIR", "PTL_WIN")) else "u32"
        out.append(f"pub const {k}: {ty} = {v};\n")
    out.append("\n// ---------------------------------------------------------------- structs\n")
    for name, fields in m["structs"].items():
        if not fields:
            out.append(f"#[repr(C)]\npub struct {name} {{\n    _private: [u8; 0],\n}}\n\n")
            continue
        out.append(f"#[repr(C)]\n#[derive(Clone, Copy)]\npub struct {name} {{\n")
        for f, t in fields:
            out.append(f"    pub {ffi_parse.rust_ident(f)}: {t},\n")
        out.append("}\n\n")
    out.append("// ---------------------------------------------------------------- functions\n")
    out.append('#[link(name = "portello_b200")]\nextern "C" {\n')
    for name, (ret, args) in m["functions"].items():
        a = ", ".join(f"{ffi_parse.rust_ident(n)}: {t}" for n, t in args)
        out.append(f"    pub fn {name}({a})" + ("" if ret == "()" else f" -> {ret}") + ";\n")
    out.append("}\n")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write("".join(out))
    print(f"wrote {OUT}: {len(m['structs'])} structs, {len(m['functions'])} functions, {len(m['consts'])} constants")


if __name__ == "__main__":
    main()
