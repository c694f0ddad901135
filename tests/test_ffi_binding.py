"""The drop-in boundary cannot rot: include/portello_b200.h (the contract), portello-b200-sys/src/lib.rs (the binding a
portello maintainer would use; Rust is absent from this image, so it ships as source) and the ctypes mirrors the tests
and bench.py use (portello_b200/abi.py, lib.py) must agree on every struct (field names, order, types) and every function
(argument names, order, types, return type), and the built library must export every declared symbol."""
import ctypes as C
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ffi_parse  # noqa: E402

HEADER = os.path.join(ROOT, "include", "portello_b200.h")
RUST = os.path.join(ROOT, "portello-b200-sys", "src", "lib.rs")


@pytest.fixture(scope="module")
def models():
    return ffi_parse.parse_header(HEADER), ffi_parse.parse_rust(RUST)


def test_type_translation():
    t = ffi_parse.c_type_to_rust
    assert t("const uint8_t* const*") == "*const *const u8"
    assert t("ptl_ctx**") == "*mut *mut ptl_ctx"
    assert t("const char* const*") == "*const *const c_char"
    assert t("const char**") == "*mut *const c_char"
    assert t("uint64_t") == "u64" and t("int") == "c_int" and t("void*") == "*mut c_void" and t("const ptl_ctx*") == "*const ptl_ctx"


def test_header_is_parsed_completely(models):
    h, _ = models
    src = ffi_parse.strip_c_comments(open(HEADER).read())
    import re
    declared = set(re.findall(r"\b(ptl_[a-z0-9_]+)\s*\(", src))
    assert declared == set(h["functions"]), declared ^ set(h["functions"])
    assert {"ptl_batch", "ptl_result", "ptl_contig_segments", "ptl_contig_records", "ptl_read_records", "ptl_read_extras"} <= set(h["structs"])
    assert len(h["structs"]["ptl_batch"]) == 20 and [f for f, _ in h["structs"]["ptl_batch"]][-3:] == ["indel_win", "rseg_win_begin", "n_indel_win"]  # incl. indel_win, rseg_win_begin, n_indel_win (the documented binding had lost them)


def test_rust_structs_match_header(models):
    h, r = models
    assert set(h["structs"]) == set(r["structs"]), set(h["structs"]) ^ set(r["structs"])
    for name, fields in h["structs"].items():
        want = [(ffi_parse.rust_ident(f), t) for f, t in fields]
        assert r["structs"][name] == want, f"{name}: header {want} vs rust {r['structs'][name]}"


def test_rust_functions_match_header(models):
    h, r = models
    assert set(h["functions"]) == set(r["functions"]), set(h["functions"]) ^ set(r["functions"])
    for name, (ret, args) in h["functions"].items():
        want = (ret, [(ffi_parse.rust_ident(n), t) for n, t in args])
        assert r["functions"][name] == want, f"{name}: header {want} vs rust {r['functions'][name]}"


def test_rust_constants_match_header(models):
    h, r = models
    assert h["consts"] == r["consts"]


RUST_TO_CTYPES = {"u8": C.c_uint8, "u16": C.c_uint16, "u32": C.c_uint32, "u64": C.c_uint64, "i8": C.c_int8, "i16": C.c_int16, "i32": C.c_int32,
                  "i64": C.c_int64, "c_int": C.c_int, "f32": C.c_float, "f64": C.c_double, "usize": C.c_size_t}


def test_ctypes_mirrors_match_header(models):
    """Every ctypes Structure of the Python harness that mirrors a header struct: same field names in the same order, and
    the same width / pointer-ness per field."""
    h, _ = models
    from portello_b200 import abi, lib
    mirrors = {"ptl_contig_segments": abi.ContigSegmentsC, "ptl_contig_records": abi.ContigRecordsC, "ptl_batch": abi.BatchC,
               "ptl_result": abi.ResultC, "ptl_read_quals": abi.ReadQualsC, "ptl_record_bases": abi.RecordBasesC,
               "ptl_read_extras": abi.ReadExtrasC, "ptl_bgzf_stream": abi.BgzfStreamC, "ptl_bam_records": abi.BamRecordsC,
               "ptl_split_segments": abi.SplitSegmentsC, "ptl_read_records": lib.ReadRecordsC}
    for name, cls in mirrors.items():
        want = h["structs"][name]
        got = [(f[0], f[1]) for f in cls._fields_]
        assert [f for f, _ in got] == [f for f, _ in want], f"{name}: field order {[f for f, _ in got]} vs header {[f for f, _ in want]}"
        for (f, ct), (_, rt) in zip(got, want):
            if rt.startswith("*"):
                assert C.sizeof(ct) == C.sizeof(C.c_void_p), f"{name}.{f}: header has a pointer"
            elif rt in RUST_TO_CTYPES:
                assert C.sizeof(ct) == C.sizeof(RUST_TO_CTYPES[rt]) and not hasattr(ct, "contents"), f"{name}.{f}: {ct} vs {rt}"
            else:
                assert C.sizeof(ct) == C.sizeof(mirrors[rt]), f"{name}.{f}: nested {rt}"


def test_library_exports_every_declared_symbol(models):
    h, _ = models
    from portello_b200 import lib
    dll = lib.load().dll
    missing = [f for f in h["functions"] if not hasattr(dll, f)]
    assert not missing, missing


def test_generated_file_is_current():
    """tools/gen_rust_sys.py is deterministic: regenerating must reproduce the committed file."""
    import subprocess
    before = open(RUST).read()
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py")], check=True, capture_output=True)
    assert open(RUST).read() == before, "portello-b200-sys/src/lib.rs is stale: run python tools/gen_rust_sys.py"
