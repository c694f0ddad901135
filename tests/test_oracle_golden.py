"""Pins the CPU oracle to the reference's own unit-test vectors (tests/golden/reference_unit_vectors.json,
transcribed from the Rust in-file tests; each vector cites its file:line)."""
import json
import os

import pytest

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_unit_vectors.json")))


def ids(key):
    return [f"{key}{i}" for i in range(len(G[key]))]


@pytest.mark.parametrize("v", G["liftover"], ids=ids("liftover"))
def test_liftover(oracle, v):
    got = oracle.liftover(v["c2r"], v["c2r_pos"], v["pos"], v["cigar"])
    exp = v["expect"]
    assert got == (None if exp is None else (exp["pos"], exp["cigar"])), v["src"]


@pytest.mark.parametrize("v", G["simplify"], ids=ids("simplify"))
def test_simplify(oracle, v):
    assert oracle.simplify(v["pos"], v["cigar"], v["ref"], v["read"]) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("v", G["shift"], ids=ids("shift"))
def test_shift(oracle, v):
    assert oracle.shift(v["dir"], v["pos"], v["cigar"], v["ref"], v["read"]) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("v", G["homology"], ids=ids("homology"))
def test_homology(oracle, v):
    rng, hom = oracle.homology(v["ref"], v["ref_range"], v["read"], v["read_range"])
    assert rng == v["expect"]["range"] and hom == v["expect"]["hom"], v["src"]


def test_tree_map(oracle):
    v = G["tree_map"][0]
    assert oracle.tree_map_ref_pos(v["ref_pos"], v["cigar"], v["ignore_hard_clip"], 4) == v["ref_pos_of"]
    rr = v["ref_range"]
    assert oracle.tree_map_ref_range(v["ref_pos"], v["cigar"], v["ignore_hard_clip"], rr["start"], rr["end"]) == rr["expect"]
    v = G["tree_map"][1]
    assert oracle.tree_map(v["ref_pos"], v["cigar"], v["ignore_hard_clip"]) == v["map"]


def test_cigar_utils(oracle):
    for v in G["compress"]:
        assert oracle.compress(v["cigar"]) == v["expect"], v["src"]
    for v in G["cleanup"]:
        assert oracle.cleanup(v["cigar"]) == (v["expect"]["shift"], v["expect"]["cigar"]), v["src"]
    for v in G["read_clip_positions"]:
        assert oracle.read_clip_positions(v["cigar"], v["ignore_hard_clip"]) == v["expect"], v["src"]
    for v in G["offsets"]:
        r, q = oracle.offsets(v["cigar"], v["ref_pos"], v["read_pos"], v["ignore_hard_clip"])
        assert q == v["expect_read"], v["src"]
        if v["expect_ref"] is not None:
            assert r == v["expect_ref"], v["src"]
    for v in G["strip_clip"]:
        assert oracle.strip_clip(v["trailing"], v["cigar"]) == v["expect"], v["src"]
    for v in G["alignment_end"]:
        r, _ = oracle.offsets(v["cigar"], v["pos"], 0, False)
        assert r[-1] == v["expect"], v["src"]
    for v in G["clip_read_edges"]:
        assert oracle.clip_read_edges(v["cigar"], v["left"], v["right"]) == (v["expect"]["cigar"], v["expect"]["ref_shift"]), v["src"]


@pytest.mark.parametrize("v", G["clip_seg_isec_range"], ids=ids("clip_seg_isec_range"))
def test_clip_seg_isec_range(oracle, v):
    got = oracle.clip_seg_isec_range(v["so"], v["pos"], v["is_fwd"], v["cigar"], v["isec"])
    assert not got["eliminated"]
    assert (got["pos"], got["cigar"], got["so"]) == (v["expect"]["pos"], v["expect"]["cigar"], v["expect"]["so"]), v["src"]


def test_split_segments_and_sa(oracle):
    for v in G["split_segments"]:
        assert oracle.split_segments(v["names"], v["tid"], v["pos"], v["flag"], v["mapq"], v["cigar"], v["sa"]) == v["expect"], v["src"]
    for v in G["sa_parse"]:
        segs = oracle.parse_sa(v["sa"])
        e = v["expect"]
        assert len(segs) == e["count"] and segs[2]["rname"] == e["rname_2"] and segs[1]["pos"] == e["pos_1"] and segs[2]["is_fwd"] == e["is_fwd_2"]


def test_misc(oracle):
    for v in G["region_segments"]:
        assert oracle.region_segments(v["size"], v["segment_size"]) == v["expect"], v["src"]
    for v in G["rev_comp"]:
        assert oracle.rev_comp(v["seq"]) == v["expect"], v["src"]


def test_reg2bin_matches_sam_spec(oracle):
    # SAM spec §5.3 reg2bin (the function htslib implements); bam_reg2bin must agree (lib/rust-vc-utils/src/bam_utils/util.rs:10-35)
    import random

    def reg2bin(beg, end):
        end -= 1
        if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
        if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
        if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
        if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
        if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
        return 0

    rnd = random.Random(7)
    for _ in range(20000):
        b = rnd.randrange(0, 1 << 29)
        e = b + rnd.choice([1, 2, 100, 15000, 1 << 14, 1 << 17, 1 << 20, rnd.randrange(1, 1 << 22)])
        e = min(e, (1 << 29))
        if e <= b:
            continue
        assert oracle.reg2bin(b, e) == reg2bin(b, e)


def test_gci_derived(oracle):
    # NOT a reference vector (the _no_align_match variant is untested upstream): arithmetic restated from
    # score_alignment.rs:68-74,138-165: '=' bases / ('=' bases + X bases + #I + #D + #N events); M is an error.
    assert oracle.gci("10=2X5=1I3D4=") == pytest.approx(19 / (19 + 2 + 1 + 1))
    assert oracle.gci("5S5H") == 1.0
    assert oracle.gci("5M") == -1.0
