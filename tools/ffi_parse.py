"""Parsers of the two sides of the drop-in boundary: the C header (include/portello_b200.h) and the Rust `-sys` crate
(portello-b200-sys/src/lib.rs).  Both are reduced to the same canonical model

    structs   {name: [(field, type), ...]}          (opaque structs have no fields)
    functions {name: (return type, [(arg, type), ...])}
    enums     {name: value}                          (anonymous C enums / #define constants  <->  Rust `pub const`)

with types written the Rust way (`*const u8`, `*mut *mut ptl_ctx`, `c_int`, ...).  tests/test_ffi_binding.py compares the
two models field for field; tools/gen_rust_sys.py writes the Rust file from the header model.
"""
from __future__ import annotations

import re

C_SCALARS = {
    "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int8_t": "i8", "int16_t": "i16", "int32_t": "i32",
    "int64_t": "i64", "int": "c_int", "char": "c_char", "float": "f32", "double": "f64", "size_t": "usize", "void": "c_void",
}


def strip_c_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def c_type_to_rust(decl: str) -> str:
    """`const uint8_t* const*` -> `*const *const u8`; `ptl_ctx**` -> `*mut *mut ptl_ctx`; `uint32_t` -> `u32`."""
    toks = re.findall(r"[A-Za-z_][A-Za-z_0-9]*|\*", decl)
    toks = [t for t in toks if t != "struct"]
    # the base type is the first identifier that is not `const`; a `const` directly around it qualifies the pointee of the
    # first `*`; a `const` after a `*` qualifies the pointee of the NEXT `*`
    base = next(t for t in toks if t not in ("const", "*"))
    i = toks.index(base)
    pointee_const = "const" in toks[:i] or (i + 1 < len(toks) and toks[i + 1] == "const")
    rest = toks[i + 1:]
    if rest and rest[0] == "const":
        rest = rest[1:]
    out = C_SCALARS.get(base, base)
    k = 0
    while k < len(rest):
        assert rest[k] == "*", decl
        out = ("*const " if pointee_const else "*mut ") + out
        pointee_const = k + 1 < len(rest) and rest[k + 1] == "const"
        k += 2 if pointee_const else 1
    return out


def _split_decl(decl: str):
    """`const uint64_t* contig_len` -> (name, rust type)."""
    decl = decl.strip()
    m = re.match(r"^(.*?)([A-Za-z_][A-Za-z_0-9]*)\s*(\[\s*\d*\s*\])?$", decl, flags=re.S)
    assert m, decl
    return m.group(2), c_type_to_rust(m.group(1))


def parse_header(path: str):
    src = strip_c_comments(open(path).read())
    structs, functions, consts = {}, {}, {}
    # #define NAME value
    for m in re.finditer(r"^[ \t]*#define\s+(PTL_[A-Z0-9_]+)\s+([0-9a-fxA-FXu]+)\s*$", src, flags=re.M):
        consts[m.group(1)] = int(m.group(2).rstrip("uU"), 0)
    src = re.sub(r"^[ \t]*#.*$", " ", src, flags=re.M)
    src = src.replace('extern "C" {', " ")
    # enums
    for m in re.finditer(r"enum\s*\{(.*?)\}\s*;", src, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = [x.strip() for x in item.split("=")]
                nxt = int(v, 0)
            else:
                k = item
            consts[k] = nxt
            nxt += 1
    src = re.sub(r"enum\s*\{.*?\}\s*;", " ", src, flags=re.S)
    # opaque structs
    for m in re.finditer(r"typedef\s+struct\s+([A-Za-z_0-9]+)\s+([A-Za-z_0-9]+)\s*;", src):
        structs.setdefault(m.group(2), [])
    src = re.sub(r"typedef\s+struct\s+[A-Za-z_0-9]+\s+[A-Za-z_0-9]+\s*;", " ", src)
    # struct definitions
    for m in re.finditer(r"typedef\s+struct\s*([A-Za-z_0-9]*)\s*\{(.*?)\}\s*([A-Za-z_0-9]+)\s*;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            if decl.strip():
                fields.append(_split_decl(decl))
        structs[m.group(3)] = fields
    src = re.sub(r"typedef\s+struct\s*[A-Za-z_0-9]*\s*\{.*?\}\s*[A-Za-z_0-9]+\s*;", " ", src, flags=re.S)
    # functions
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9\s\*]*?)\b(ptl_[a-z0-9_]+)\s*\(([^()]*(?:\([^()]*\)[^()]*)*)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if name in structs:
            continue
        arglist = []
        if args and args != "void":
            for a in args.split(","):
                arglist.append(_split_decl(a))
        functions[name] = ("()" if ret == "void" else c_type_to_rust(ret), arglist)
    return {"structs": structs, "functions": functions, "consts": consts}


def parse_rust(path: str):
    src = open(path).read()
    src = re.sub(r"//[^\n]*", " ", src)
    structs, functions, consts = {}, {}, {}
    for m in re.finditer(r"pub\s+const\s+([A-Z0-9_]+)\s*:\s*[a-z0-9_]+\s*=\s*(-?[0-9a-fx_]+)\s*;", src):
        consts[m.group(1)] = int(m.group(2).replace("_", ""), 0)
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[[^\]]*\]\s*)*pub\s+struct\s+([A-Za-z_0-9]+)\s*\{(.*?)\}", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(","):
            decl = decl.strip()
            if not decl:
                continue
            fm = re.match(r"^(pub\s+)?([A-Za-z_][A-Za-z_0-9]*)\s*:\s*(.+)$", decl, flags=re.S)
            assert fm, decl
            if fm.group(2) == "_private":
                continue  # opaque marker
            fields.append((fm.group(2), " ".join(fm.group(3).split())))
        structs[m.group(1)] = fields
    for m in re.finditer(r"pub\s+fn\s+(ptl_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*([^;]+?))?\s*;", src, flags=re.S):
        args = []
        for a in m.group(2).split(","):
            a = a.strip()
            if a:
                n, t = a.split(":", 1)
                args.append((n.strip(), " ".join(t.split())))
        functions[m.group(1)] = (" ".join(m.group(3).split()) if m.group(3) else "()", args)
    return {"structs": structs, "functions": functions, "consts": consts}


RUST_KEYWORDS = {"in", "ref", "type", "fn", "mod", "move", "match", "loop", "impl", "self", "use", "where", "box", "final", "override"}


def rust_ident(name: str) -> str:
    return name + "_" if name in RUST_KEYWORDS else name
