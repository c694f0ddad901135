"""Replays the reference's own unit-test vectors (tests/golden) through the CUDA kernels via the C-ABI, one stage at a
time (ptl_lift_submit_ex stage masks), and checks the same calls against the oracle."""
import json
import os

import numpy as np
import pytest

import helpers
from portello_b200 import abi, lib

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_unit_vectors.json")))


@pytest.fixture(scope="module")
def gpu():
    ctx = lib.GpuContext(0, 1)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("long_ops", [64, 0], ids=["thread", "warp"])
@pytest.mark.parametrize("v", G["liftover"], ids=[f"liftover{i}" for i in range(len(G["liftover"]))])
def test_liftover_vectors(gpu, v, long_ops):
    gpu.set_long_pair_ops(long_ops)  # 0: the warp-cooperative liftover (lift_warp.cuh) instead of the per-thread walk
    c2r = v["c2r"] if v["c2r"] is not None else "100S"  # an empty map: a segment whose CIGAR has no aligned run
    contig_len = helpers.cigar_read_len(c2r)
    segs, batch = helpers.single_pair_case(c2r, v["c2r_pos"], True, contig_len, None, v["pos"], v["cigar"], [], helpers.cigar_read_len(v["cigar"]))
    gpu.set_contig_segments(segs)
    res = gpu.lift(batch, stage_mask=abi.STAGE_LIFTOVER)
    if v["expect"] is None:
        assert res.n_records == 0 and res.n_pairs == 1 and res.n_lifted == 0, v["src"]
    else:
        assert res.n_records == 1, v["src"]
        assert (int(res.rec_pos[0]), res.record_cigar(0)) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("v", G["simplify"], ids=[f"simplify{i}" for i in range(len(G["simplify"]))])
def test_simplify_vectors(gpu, v):
    ref = helpers.relabel(v["ref"])
    read = helpers.relabel(v["read"])
    gpu.set_reference([np.frombuffer(ref.encode(), dtype=np.uint8)])
    segs, batch = helpers.single_pair_case(f"{len(ref)}=", 0, True, len(ref), None, v["pos"], v["cigar"], abi.pack_seq4(read), len(read))
    gpu.set_contig_segments(segs)
    res = gpu.lift(batch, stage_mask=abi.STAGE_SIMPLIFY)
    assert res.n_records == 1
    assert (int(res.rec_pos[0]), res.record_cigar(0)) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("long_ops", [64, 0], ids=["thread", "warp"])
@pytest.mark.parametrize("v", [x for x in G["shift"] if x["dir"] == "left"], ids=lambda v: v["cigar"])
def test_left_shift_vectors(gpu, v, long_ops):
    gpu.set_long_pair_ops(long_ops)  # 0: warp_left_shift (lift_warp.cuh) instead of the per-thread builder
    # The kernel applies the shift on the contig's reverse strand: feed the mirrored alignment so that the kernel's own
    # reversal (rev_pos, reversed CIGAR; src/read_alignment_scanner.rs:163-167) reconstructs the vector's input.
    ref = helpers.relabel(v["ref"])
    read = helpers.relabel(v["read"])
    cig = abi.cigar_from_string(v["cigar"])
    contig_len = len(ref)
    ref_len = helpers.cigar_ref_len(v["cigar"])
    pos_fwd = contig_len - (v["pos"] + ref_len)
    segs, batch = helpers.single_pair_case(f"{contig_len}=", 0, False, contig_len, ref, pos_fwd, cig[::-1].copy(), abi.pack_seq4(read), len(read),
                                           read_flag=0, rseg_fwd=0)  # need_flip = (!contig_fwd) ^ (rev == seg_fwd) = 1 ^ 1 = 0
    gpu.set_contig_segments(segs)
    res = gpu.lift(batch, stage_mask=abi.STAGE_LEFT_SHIFT)
    assert res.n_records == 1 and int(res.rec_need_flip[0]) == 0
    assert (int(res.rec_pos[0]), res.record_cigar(0)) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


def test_tree_map_vectors(gpu, oracle):
    for v in G["tree_map"]:
        if v["ignore_hard_clip"]:
            continue  # portello always builds its tables with ignore_hard_clip = false
        cl = helpers.cigar_read_len(v["cigar"])
        segs, _ = helpers.single_pair_case(v["cigar"], v["ref_pos"], True, cl, None, 0, "1M", [], 1)
        gpu.set_contig_segments(segs)
        keys, vals = gpu.get_segment_table(0)
        got = [[int(k), (None if x < 0 else int(x))] for k, x in zip(keys, vals)]
        assert got == v["map"], v["src"]
    # the reference's own vector uses ignore_hard_clip=true; with false the same CIGAR shifts keys by the 2H
    segs, _ = helpers.single_pair_case("2H2M1I1M", 9, True, 6, None, 0, "1M", [], 1)
    gpu.set_contig_segments(segs)
    keys, vals = gpu.get_segment_table(0)
    assert [list(map(int, keys)), list(map(int, vals))] == [[2, 4, 5, 6], [9, -1, 11, -1]]
    assert oracle.tree_map(9, "2H2M1I1M", False) == [[2, 9], [4, None], [5, 11], [6, None]]


@pytest.mark.parametrize("n_indels,read_cigar", [
    (40, "3S5000=2I1000=3D2500="),       # one op crossing > 64 table keys: several piece rounds, a `big` window
    (120, "10=1X4000=5D30=7I4000=20S"),
    (33, "2000=1I2000="),
], ids=["40", "120", "33"])
def test_warp_liftover_many_keys_per_op(oracle, n_indels, read_cigar):
    c2r = "".join(f"{40 + (k % 7)}={1 + k % 3}{'ID'[k % 2]}" for k in range(n_indels)) + "9000="
    contig_len = helpers.cigar_read_len(c2r)
    rlen = helpers.cigar_read_len(read_cigar)
    ref = np.frombuffer(b"ACGT" * ((helpers.cigar_ref_len(c2r) + 200) // 4 + 1), dtype=np.uint8).copy()
    out = []
    for L, long_ops in ((oracle, 64), (lib.load(), 0), (lib.load(), 1 << 30)):
        segs, batch = helpers.single_pair_case(c2r, 100, True, contig_len, None, 17, read_cigar, [], rlen)
        ctx = abi.Context(L, 0, 1)
        ctx.set_long_pair_ops(long_ops)
        ctx.set_reference([ref])
        ctx.set_contig_segments(segs)
        r = ctx.lift(batch, stage_mask=abi.STAGE_LIFTOVER)
        assert r.n_records == 1
        out.append((int(r.rec_pos[0]), r.record_cigar(0)))
        ctx.close()
    assert out[0] == out[1] == out[2], out
