"""The composition itself, pinned from its DEFINITION rather than from a restatement of the reference's algorithm
(a4 / a6 had no vector that the builder had not derived from the same reading of the code).

A lifted alignment says, for every read base, which reference base it sits on.  That is decided by the two input
alignments alone: read base q sits on contig base c (read->contig CIGAR), contig base c sits on reference base r
(contig->reference CIGAR), so q sits on r -- or on nothing, when c is an insertion / clip of the contig alignment.  The
model below composes the two per-base maps with dictionaries and knows nothing about tables, blocks, pieces or sinks; the
liftover stage of every implementation (oracle, the device code compiled for the host, the GPU) must produce a CIGAR
that induces exactly that map, on forward- and reverse-strand contig segments (for the latter the output is the reverse
complement of the record: base q of the record is base len-1-q of the output, contig base c is base L-1-c of the strand
the segment's CIGAR is written on; src/read_alignment_scanner.rs:153-183).  Only the liftover stage runs (no left shift, no
simplify: those move indels inside repeats, which changes the map without changing the alignment score)."""
import numpy as np
import pytest

import helpers
import oracle_lib
from portello_b200 import abi

QUERY_OPS, REF_OPS, MATCH_OPS = (0, 1, 4, 5, 7, 8), (0, 2, 3, 7, 8), (0, 7, 8)


def aligned_pairs(ops, pos):
    """(query index, target position) of every aligned base of a CIGAR; and the query length."""
    q, t, out = 0, pos, []
    for x in ops:
        op, l = int(x) & 15, int(x) >> 4
        if op in MATCH_OPS:
            out += [(q + i, t + i) for i in range(l)]
        q += l if op in QUERY_OPS else 0
        t += l if op in REF_OPS else 0
    return out, q


def compose(read_ops, read_pos, c2r_ops, c2r_pos, c2r_fwd, contig_len):
    q2c, qlen = aligned_pairs(read_ops, read_pos)
    c2r = dict(aligned_pairs(c2r_ops, c2r_pos)[0])
    if c2r_fwd:
        want = [(q, c2r[c]) for q, c in q2c if c in c2r]
    else:
        want = sorted((qlen - 1 - q, c2r[contig_len - 1 - c]) for q, c in q2c if contig_len - 1 - c in c2r)
    return want, qlen


def random_ops(rng, n_blocks, lead_clip, trail_clip, match_ops=(0, 7, 8)):
    ops = []
    if lead_clip:
        ops.append((int(rng.choice([4, 5])), int(rng.integers(1, 40))))
    for b in range(n_blocks):
        ops.append((int(rng.choice(match_ops)), int(rng.integers(1, 60))))
        if b + 1 < n_blocks:
            kind = rng.random()
            if kind < 0.4:
                ops.append((1, int(rng.integers(1, 12))))
            elif kind < 0.8:
                ops.append((int(rng.choice([2, 2, 2, 3])), int(rng.integers(1, 12))))
            elif kind < 0.9:
                ops += [(1, int(rng.integers(1, 6))), (2, int(rng.integers(1, 6)))]
            # else: two match ops of (possibly) different kinds next to each other
    if trail_clip:
        ops.append((int(rng.choice([4, 5])), int(rng.integers(1, 40))))
    return np.array([(l << 4) | o for o, l in ops], np.uint32)


def cases(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        c2r = random_ops(rng, int(rng.integers(1, 25)), rng.random() < 0.5, rng.random() < 0.5)
        contig_len = sum(int(x) >> 4 for x in c2r if (int(x) & 15) in QUERY_OPS)
        read = random_ops(rng, int(rng.integers(1, 12)), rng.random() < 0.4, rng.random() < 0.4)
        span = sum(int(x) >> 4 for x in read if (int(x) & 15) in REF_OPS)
        if span >= contig_len:
            continue
        out.append(dict(c2r=c2r, c2r_pos=int(rng.integers(0, 1000)), c2r_fwd=bool(rng.random() < 0.5), contig_len=contig_len,
                        read=read, read_pos=int(rng.integers(0, contig_len - span + 1))))
    return out


def check(make_ctx, seed, n=150, long_ops=None):
    n_lifted = n_rev = n_unaligned = 0
    for v in cases(seed, n):
        qlen = sum(int(x) >> 4 for x in v["read"] if (int(x) & 15) in QUERY_OPS)
        ref_len = v["c2r_pos"] + sum(int(x) >> 4 for x in v["c2r"] if (int(x) & 15) in REF_OPS) + 10
        rev_seq = None if v["c2r_fwd"] else b"A" * v["contig_len"]
        segs, batch = helpers.single_pair_case(v["c2r"], v["c2r_pos"], v["c2r_fwd"], v["contig_len"], rev_seq, v["read_pos"], v["read"], [], qlen)
        ctx = make_ctx()
        if long_ops is not None:
            ctx.set_long_pair_ops(long_ops)
        ctx.set_reference([np.full(ref_len, ord("A"), np.uint8)])
        ctx.set_contig_segments(segs)
        res = ctx.lift(batch, stage_mask=abi.STAGE_LIFTOVER)
        want, _ = compose(v["read"], v["read_pos"], v["c2r"], v["c2r_pos"], v["c2r_fwd"], v["contig_len"])
        if not want:  # no read base reaches the reference: liftover returns None, nothing is lifted
            assert res.n_lifted == 0 and not any(int(s) == 1 for s in res.rec_status), v
            n_unaligned += 1
            continue
        assert res.n_records == 1 and res.rec_status[0] == 1, v
        k0, k1 = int(res.rec_cigar_begin[0]), int(res.rec_cigar_begin[1])
        got, got_qlen = aligned_pairs(res.cigar[k0:k1], int(res.rec_pos[0]))
        assert got_qlen == qlen, (v, res.record_cigar(0))
        assert got == want, (v, res.record_cigar(0), [p for p in zip(got, want) if p[0] != p[1]][:3])
        assert int(res.rec_pos[0]) == want[0][1]            # the record starts on its first aligned base (leading deletions are dropped)
        ops = [int(x) & 15 for x in res.cigar[k0:k1]]
        first_m, last_m = min(i for i, o in enumerate(ops) if o == 0), max(i for i, o in enumerate(ops) if o == 0)
        assert all(o not in (1, 2) for o in ops[:first_m] + ops[last_m + 1:]), res.record_cigar(0)   # clean edges: no I / D outside the matches (cigar/mod.rs:233-291)
        assert all(a != b or a == 6 for a, b in zip(ops, ops[1:])) and all(int(x) >> 4 for x in res.cigar[k0:k1])  # compressed (:204-228)
        n_lifted += 1
        n_rev += not v["c2r_fwd"]
    assert n_lifted > n * 0.6 and n_rev > n * 0.2
    return n_unaligned


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_oracle_liftover_induces_the_composed_base_map(seed):
    check(lambda: abi.Context(oracle_lib.load(), 0, 1), seed)


@pytest.mark.parametrize("long_ops", [None, 0], ids=["thread-per-pair", "warp-per-pair"])
def test_device_code_liftover_induces_the_composed_base_map(long_ops):
    import emul_lib
    check(lambda: abi.Context(emul_lib.load(), 0, 1), 21, n=120, long_ops=long_ops)


@pytest.mark.gpu
@pytest.mark.parametrize("long_ops", [None, 0], ids=["thread-per-pair", "warp-per-pair"])
def test_gpu_liftover_induces_the_composed_base_map(long_ops):
    from portello_b200 import lib
    check(lambda: lib.GpuContext(0, 1), 31, n=150, long_ops=long_ops)


# ------------------------------------------------------------------------------------------------------------------------
# a11 / a12 (contig preparation: record assembly, repeated-match trimming, colinear joining) from what it may and may not do

def _pair_keys(ops, pos, contig_len, reverse, chrom):
    """One 64-bit key per aligned (contig base, reference base) of a CIGAR, contig base in FORWARD contig coordinates."""
    keys = []
    q, t = 0, int(pos)
    for x in ops:
        op, l = int(x) & 15, int(x) >> 4
        if op in MATCH_OPS and l:
            qi = np.arange(q, q + l, dtype=np.int64)
            c = (contig_len - 1 - qi) if reverse else qi
            keys.append((c << 34) | ((np.arange(t, t + l, dtype=np.int64) + (int(chrom) << 28)) << 1) | int(bool(reverse)))
        q += l if op in QUERY_OPS else 0
        t += l if op in REF_OPS else 0
    return (np.concatenate(keys) if keys else np.zeros(0, np.int64)), q


@pytest.mark.parametrize("kw", [dict(seed=5), dict(seed=11, chrom_len=3_000_000, contigs_per_chrom=3, junction_per_mb=8),
                                dict(seed=12, chrom_len=6_000_000, contigs_per_chrom=2, junction_per_mb=6, rev_contig_frac=1.0),
                                dict(seed=14, chrom_len=4_000_000, contigs_per_chrom=3, junction_per_mb=10, sv_per_mb=20.0)])
def test_contig_preparation_never_invents_an_alignment(kw):
    """Trimming may only take aligned bases away (clip_alignment_read_edges), joining may only put an insertion and a
    deletion between two segments (contig_colinear_segment_joiner.rs:57-122): every (contig base -> reference base, strand)
    of the prepared segments must be one that some input record of that contig states, every prepared CIGAR must span
    the whole contig, and what trimming removes is exactly the double cover -- afterwards no contig base is aligned twice."""
    from portello_b200 import lib, synth
    s = synth.make("tiny", n_reads=10, **kw)
    raw, got = s.contig_records, lib.load().prepare_contig_records(s.contig_records)
    raw_cig = np.ctypeslib.as_array(raw.cigar, (int(raw.cigar_begin[raw.n_records]),))
    n_seg_total = n_pairs_total = n_joined_gaps = 0
    for ctg in range(raw.n_contigs):
        L = int(raw.contig_len[ctg])
        stated = []
        for k in range(raw.n_records):
            if raw.contig_id[k] != ctg or (raw.flag[k] & 0x104):     # other contig / unmapped / secondary
                continue
            keys, qlen = _pair_keys(raw_cig[int(raw.cigar_begin[k]): int(raw.cigar_begin[k + 1])], raw.pos[k], L, bool(raw.flag[k] & 0x10), raw.tid[k])
            assert qlen == L
            stated.append(keys)
        stated = np.sort(np.concatenate(stated)) if stated else np.zeros(0, np.int64)

        def is_stated(keys):  # (sorted search: numpy's hash-based unique / isin take seconds per call on millions of keys)
            at = np.minimum(np.searchsorted(stated, keys), max(len(stated) - 1, 0))
            return len(stated) > 0 and bool(np.all(stated[at] == keys))

        mine = []
        for g in range(int(got.contig_seg_begin[ctg]), int(got.contig_seg_begin[ctg + 1])):
            ops = got.cigar[int(got.seg_cigar_begin[g]): int(got.seg_cigar_begin[g + 1])]
            keys, qlen = _pair_keys(ops, got.seg_pos[g], L, not got.seg_is_fwd[g], got.seg_chrom_index[g])
            assert qlen == L, (ctg, g)
            assert is_stated(keys), (ctg, g)
            # the segment's seq-order interval is where its aligned bases are (in sequencing order = forward contig coordinates)
            c = keys >> 34
            assert int(got.seg_seq_order_start[g]) <= int(c.min()) and int(c.max()) < int(got.seg_seq_order_end[g])
            mine.append(keys)
            n_seg_total += 1
            n_joined_gaps += sum(1 for a, b in zip(ops, ops[1:]) if (int(a) & 15, int(b) & 15) == (1, 2) and (int(b) >> 4) > 50)
        if mine:
            allk = np.concatenate(mine)
            contig_base = allk >> 34
            assert np.all(np.diff(np.sort(contig_base)) > 0), ctg     # no contig base is aligned twice any more
            n_pairs_total += len(allk)
    assert n_seg_total >= raw.n_contigs - 1 and n_pairs_total > 0.5 * int(np.sum(np.ctypeslib.as_array(raw.contig_len, (raw.n_contigs,))))


# ------------------------------------------------------------------------------------------------------------------------
# a5 (left_shift_indels) from what a left shift is: the same alignment columns, indels as far left as the sequences allow

def _columns(ops, pos, read, ref):
    """Walk an alignment: (number of mismatching match columns, total I, total D, read length, reference end)."""
    q, t, mism, ins, dele = 0, pos, 0, 0, 0
    for x in ops:
        op, l = int(x) & 15, int(x) >> 4
        if op in MATCH_OPS:
            mism += sum(1 for i in range(l) if read[q + i] != ref[t + i])
        ins += l if op == 1 else 0
        dele += l if op == 2 else 0
        q += l if op in QUERY_OPS else 0
        t += l if op in REF_OPS else 0
    return mism, ins, dele, q, t


def _shift_case(rng):
    """A reference of low complexity behind a unique anchor, and a read derived from it by known edits."""
    anchor = "".join(rng.choice(list("GT"), 24))
    body = "".join(rng.choice(list("AC"), int(rng.integers(60, 400)), p=[0.7, 0.3]))
    ref = anchor + body + anchor[::-1]
    read, ops, t = [], [], 0

    def match(n, kind=7):
        nonlocal t
        read.append(ref[t:t + n])
        ops.append((kind, n))
        t += n
    match(len(anchor))
    while t < len(anchor) + len(body) - 30:
        match(int(rng.integers(1, 25)))
        k = rng.random()
        n = int(rng.integers(1, 6))
        if k < 0.4:
            read.append("".join(rng.choice(list("AC"), n, p=[0.7, 0.3])))
            ops.append((1, n))
        elif k < 0.8:
            ops.append((2, n))
            t += n
        elif k < 0.9:     # a mismatch column
            read.append("G")
            ops.append((8, 1))
            t += 1
        else:             # an insertion right next to a deletion
            read.append("".join(rng.choice(list("AC"), n)))
            ops.append((1, n))
            ops.append((2, n + 1))
            t += n + 1
    match(len(ref) - t)
    return ref, "".join(read), np.array([(l << 4) | o for o, l in ops], np.uint32)


def _separated_shift_case(rng):
    """One indel per low-complexity run, the runs kept apart by unique anchors: clusters never meet, so every indel must end
    up as far left as its run allows."""
    ref, read, ops = [], [], []

    def both(sq, kind=7):
        ref.append(sq)
        read.append(sq)
        ops.append((kind, len(sq)))
    both("".join(rng.choice(list("GT"), 20)))
    for _ in range(int(rng.integers(2, 9))):
        run = "".join(rng.choice(list("AC"), int(rng.integers(12, 40)), p=[0.75, 0.25]))
        cut, n = int(rng.integers(4, len(run) - 4)), int(rng.integers(1, 4))
        both(run[:cut])
        if rng.random() < 0.5:
            read.append("".join(rng.choice(list("AC"), n, p=[0.75, 0.25])))
            ops.append((1, n))
            both(run[cut:])
        else:
            ref.append(run[cut:cut + n])
            ops.append((2, n))
            both(run[cut + n:])
        both("".join(rng.choice(list("GT"), 14)))
    merged = []
    for o, l in ops:    # (adjacent match ops of one kind: one op)
        if merged and merged[-1][0] == o:
            merged[-1] = (o, merged[-1][1] + l)
        elif l:
            merged.append((o, l))
    return "".join(ref), "".join(read), np.array([(l << 4) | o for o, l in merged], np.uint32)


def _left_shift(make_ctx, ref, read, pos, cig, long_ops):
    ctx = make_ctx()
    if long_ops is not None:
        ctx.set_long_pair_ops(long_ops)
    L = len(ref)
    span = sum(int(x) >> 4 for x in cig if (int(x) & 15) in REF_OPS)
    segs, batch = helpers.single_pair_case(f"{L}=", 0, False, L, ref, L - (pos + span), cig[::-1].copy(), abi.pack_seq4(read), len(read), read_flag=0, rseg_fwd=0)
    ctx.set_contig_segments(segs)
    res = ctx.lift(batch, stage_mask=abi.STAGE_LEFT_SHIFT)
    assert res.n_records == 1
    return int(res.rec_pos[0]), np.asarray(res.cigar[int(res.rec_cigar_begin[0]): int(res.rec_cigar_begin[1])], np.uint32)


def check_left_shift(make_ctx, seed, n, long_ops=None):
    rng = np.random.default_rng(seed)
    n_moved = 0
    for it in range(n):
        separated = it % 2 == 1
        ref, read, cig = _separated_shift_case(rng) if separated else _shift_case(rng)
        pos2, cig2 = _left_shift(make_ctx, ref, read, 0, cig, long_ops)
        before, after = _columns(cig, 0, read, ref), _columns(cig2, pos2, read, ref)
        # the same edits on the same bases: read length, reference end, inserted and deleted bases; and no column got worse
        assert pos2 == 0 and after[3] == before[3] == len(read) and after[4] == before[4] and after[1:3] == before[1:3], (ref, read, abi.cigar_to_string(cig))
        assert after[0] <= before[0]
        # (not idempotent, by the reference's design: two clusters that the shift makes adjacent move on as ONE cluster only
        #  in a second pass, e.g. 12=4D2=1I11= -> 8M4D1I15M -> 7M1I4D16M; so "leftmost" is asked of the cases whose indels
        #  cannot meet)
        # leftmost: an insertion or deletion behind a match block cannot move one more base (the homology walk of
        # indel_breakend_homology.rs:35-49 compares ref[ref_end - 1] with read[read_end - 1])
        q, t, ops = 0, pos2, [(int(x) & 15, int(x) >> 4) for x in cig2]
        for i, (op, l) in enumerate(ops):
            if separated and op in (1, 2):
                assert i > 0 and ops[i - 1][0] in MATCH_OPS and ops[i + 1][0] in MATCH_OPS
                ref_end, read_end = t + (l if op == 2 else 0), q + (l if op == 1 else 0)
                assert ref[ref_end - 1] != read[read_end - 1], (abi.cigar_to_string(cig2), i)
            q += l if op in QUERY_OPS else 0
            t += l if op in REF_OPS else 0
        n_moved += list(cig2) != [int(x) & ~0xf | (0 if (int(x) & 15) in MATCH_OPS else int(x) & 15) for x in cig]
    assert n_moved > n // 2     # (the inputs are built to have room to move)


@pytest.mark.parametrize("seed", [41, 42])
def test_oracle_left_shift_keeps_the_columns_and_ends_leftmost(seed):
    check_left_shift(lambda: abi.Context(oracle_lib.load(), 0, 1), seed, 60)


@pytest.mark.parametrize("long_ops", [None, 0], ids=["thread-per-pair", "warp-per-pair"])
def test_device_code_left_shift_keeps_the_columns_and_ends_leftmost(long_ops):
    import emul_lib
    check_left_shift(lambda: abi.Context(emul_lib.load(), 0, 1), 43, 40, long_ops)


@pytest.mark.gpu
@pytest.mark.parametrize("long_ops", [None, 0], ids=["thread-per-pair", "warp-per-pair"])
def test_gpu_left_shift_keeps_the_columns_and_ends_leftmost(long_ops):
    from portello_b200 import lib
    check_left_shift(lambda: lib.GpuContext(0, 1), 44, 60, long_ops)


# ------------------------------------------------------------------------------------------------------------------------
# a9 (simplify_alignment_indels) from what it is for: an I/D run that holds both kinds is trimmed of the base pairs that match
# at its ends, a leftover 1I1D is one (mis)match column, and nothing else moves (src/simplify_alignment_indels.rs:35-156)

def _simplify_case(rng):
    anchor = "".join(rng.choice(list("GT"), 16))
    ref, read, ops = [anchor], [anchor], [(7, len(anchor))]
    for _ in range(int(rng.integers(2, 14))):
        n = int(rng.integers(2, 30))
        sq = "".join(rng.choice(list("ACGT"), n))
        ref.append(sq)
        read.append(sq)
        ops.append((int(rng.choice([0, 7])), n))
        k = rng.random()
        ni, nd = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        ins, dele = "".join(rng.choice(list("AC"), ni)), "".join(rng.choice(list("AC"), nd))
        if k < 0.25:
            read.append(ins)
            ops.append((1, ni))
        elif k < 0.5:
            ref.append(dele)
            ops.append((2, nd))
        else:  # a run with both kinds, in either order, sometimes in three pieces
            read.append(ins)
            ref.append(dele)
            if k < 0.7:
                ops += [(1, ni), (2, nd)]
            elif k < 0.9:
                ops += [(2, nd), (1, ni)]
            else:
                a = int(rng.integers(0, ni + 1))
                ops += [(1, a), (2, nd), (1, ni - a)]
    ref.append(anchor)
    read.append(anchor)
    ops.append((7, len(anchor)))
    return "".join(ref), "".join(read), np.array([(l << 4) | o for o, l in ops if l], np.uint32)


def _clusters(ops, pos):
    """(ref start, read start, deleted, inserted) of every I/D run of a CIGAR."""
    q, t, out, cur = 0, pos, [], None
    for x in ops:
        op, l = int(x) & 15, int(x) >> 4
        if op in (1, 2):
            if cur is None:
                cur = [t, q, 0, 0]
            cur[2 if op == 2 else 3] += l
        elif cur is not None:
            out.append(tuple(cur))
            cur = None
        q += l if op in QUERY_OPS else 0
        t += l if op in REF_OPS else 0
    return out + ([tuple(cur)] if cur else [])


def check_simplify(make_ctx, seed, n, long_ops=None):
    rng = np.random.default_rng(seed)
    n_changed = 0
    for _ in range(n):
        ref, read, cig = _simplify_case(rng)

        def run(pos, ops):
            ctx = make_ctx()
            if long_ops is not None:
                ctx.set_long_pair_ops(long_ops)
            ctx.set_reference([np.frombuffer(ref.encode(), np.uint8)])
            segs, batch = helpers.single_pair_case(f"{len(ref)}=", 0, True, len(ref), None, pos, ops, abi.pack_seq4(read), len(read))
            ctx.set_contig_segments(segs)
            res = ctx.lift(batch, stage_mask=abi.STAGE_SIMPLIFY)
            assert res.n_records == 1
            return int(res.rec_pos[0]), np.asarray(res.cigar[int(res.rec_cigar_begin[0]): int(res.rec_cigar_begin[1])], np.uint32)

        pos2, cig2 = run(0, cig)
        before, after = _columns(cig, 0, read, ref), _columns(cig2, pos2, read, ref)
        assert pos2 == 0 and after[3] == before[3] == len(read) and after[4] == before[4] == len(ref)
        # every aligned pair of the input is still there; the new ones sit inside the input's mixed runs
        old_pairs, new_pairs = set(aligned_pairs(cig, 0)[0]), set(aligned_pairs(cig2, pos2)[0])
        assert old_pairs <= new_pairs
        mixed = [c for c in _clusters(cig, 0) if c[2] and c[3]]
        for q, t in new_pairs - old_pairs:
            assert any(c[1] <= q < c[1] + c[3] and c[0] <= t < c[0] + c[2] for c in mixed)
        # what is left of a run with both kinds cannot be trimmed any further, and is not 1I1D
        for t0, q0, d, i in _clusters(cig2, pos2):
            if d and i:
                assert (d, i) != (1, 1)
                assert ref[t0 + d - 1] != read[q0 + i - 1] and ref[t0] != read[q0], abi.cigar_to_string(cig2)
        # the trimmed pairs are real matches, except the single column a 1I1D leftover turns into
        n_mismatch_new = sum(1 for q, t in new_pairs - old_pairs if read[q] != ref[t])
        assert n_mismatch_new <= len(mixed)
        pos3, cig3 = run(pos2, cig2)
        assert pos3 == pos2 and list(cig3) == list(cig2)   # idempotent
        n_changed += bool(new_pairs - old_pairs)
    assert n_changed > n // 3


@pytest.mark.parametrize("seed", [51, 52])
def test_oracle_simplify_trims_mixed_runs_and_moves_nothing_else(seed):
    check_simplify(lambda: abi.Context(oracle_lib.load(), 0, 1), seed, 60)


@pytest.mark.parametrize("long_ops", [None, 0], ids=["thread-per-pair", "warp-per-pair"])
def test_device_code_simplify_trims_mixed_runs_and_moves_nothing_else(long_ops):
    import emul_lib
    check_simplify(lambda: abi.Context(emul_lib.load(), 0, 1), 53, 40, long_ops)


@pytest.mark.gpu
@pytest.mark.parametrize("long_ops", [None, 0], ids=["thread-per-pair", "warp-per-pair"])
def test_gpu_simplify_trims_mixed_runs_and_moves_nothing_else(long_ops):
    from portello_b200 import lib
    check_simplify(lambda: lib.GpuContext(0, 1), 54, 60, long_ops)


# ------------------------------------------------------------------------------------------------------------------------
# a4 / a10 record fields from the strand algebra and the documented rules (docs/methods.md:22-24,43-45; SURVEY parity traps
# 3, 4, 12): no restatement of finish_remapped_alignment_set, only what its output must look like

def check_record_rules(ctx, s, pb, segs):
    import oracle_lib as O
    r = helpers.lift_c(ctx, pb.c, allow_panic=True)
    b = pb.c
    g = lambda p, n: np.ctypeslib.as_array(p, (n,))
    read_flag, rseg_fwd, rseg_ctg = g(b.read_flag, b.n_reads), g(b.rseg_is_fwd, b.n_read_segments), g(b.rseg_contig, b.n_read_segments)
    read_bin = g(b.read_bin, b.n_reads)
    n_multi = n_fallback = n_rev = 0
    for i in range(b.n_reads):
        k0, k1 = int(r.read_rec_begin[i]), int(r.read_rec_begin[i + 1])
        lifted = [k for k in range(k0, k1) if r.rec_status[k] == 1]
        if not lifted:
            # the unmapped fallback (:317-335): one record, unmapped, no position, MAPQ 255, the input record's bin, no CIGAR
            assert k1 - k0 == 1 and (r.rec_flag[k0] & 0x4) and not (r.rec_flag[k0] & 0x800)
            assert (int(r.rec_tid[k0]), int(r.rec_pos[k0]), int(r.rec_mapq[k0])) == (-1, -1, 255) and r.rec_bin[k0] == read_bin[i]
            assert r.rec_cigar_begin[k0] == r.rec_cigar_begin[k0 + 1]
            n_fallback += 1
            continue
        assert lifted == list(range(k0, k1))
        keys = []
        for k in lifted:
            rs, cs = int(r.rec_read_segment[k]), int(r.rec_contig_segment[k])
            assert int(b.read_seg_begin[i]) <= rs < int(b.read_seg_begin[i + 1])
            gseg = int(segs.contig_seg_begin[rseg_ctg[rs]]) + cs
            assert gseg < int(segs.contig_seg_begin[rseg_ctg[rs] + 1])
            # strands compose: the read lies reversed on the reference iff it lies on the contig and the contig on the reference
            # in different orientations; the stored bases are flipped iff that differs from the orientation of the input record
            on_ref_reverse = bool(rseg_fwd[rs]) != bool(segs.seg_is_fwd[gseg])
            assert bool(r.rec_flag[k] & 0x10) == on_ref_reverse
            assert bool(r.rec_need_flip[k]) == (on_ref_reverse != bool(read_flag[i] & 0x10))
            assert not (r.rec_flag[k] & 0x4)
            # chromosome and MAPQ are the contig segment's (trap 3: the lifted MAPQ is not the read's)
            assert int(r.rec_tid[k]) == int(segs.seg_chrom_index[gseg]) and int(r.rec_mapq[k]) == int(segs.seg_mapq[gseg])
            ops = r.cigar[int(r.rec_cigar_begin[k]): int(r.rec_cigar_begin[k + 1])]
            end = int(r.rec_pos[k]) + sum(int(x) >> 4 for x in ops if (int(x) & 15) in REF_OPS)
            assert int(r.rec_bin[k]) == O.load().reg2bin(int(r.rec_pos[k]), end)
            keys.append((rs, cs))
            n_rev += on_ref_reverse
        assert keys == sorted(keys) and len(set(keys)) == len(keys)      # (read segment in sequencing order) x (contig segment index)
        # trap 4: the primary is the FIRST record with the highest MAPQ; every other record is supplementary
        mq = [int(r.rec_mapq[k]) for k in lifted]
        prim = [k for k in lifted if not (r.rec_flag[k] & 0x800)]
        assert prim == [lifted[mq.index(max(mq))]]
        n_multi += len(lifted) > 1
    return n_multi, n_fallback, n_rev


def _record_rule_set():
    from portello_b200 import synth
    s = synth.make("tiny", seed=61, n_reads=4000, read_sa_frac=0.25, junction_per_mb=20.0, rev_contig_frac=0.5)
    return s, helpers.pack(s)


def test_oracle_and_device_code_records_follow_the_documented_rules():
    import emul_lib
    s, pb = _record_rule_set()
    for L in (oracle_lib.load(), emul_lib.load()):
        ctx = abi.Context(L, 0, 1)
        ctx.set_reference(helpers.reference_arrays(s))
        ctx.set_contig_records(s.contig_records)
        n_multi, n_fallback, n_rev = check_record_rules(ctx, s, pb, ctx.get_contig_segments())
        assert n_multi > 100 and n_fallback > 10 and n_rev > 500, (n_multi, n_fallback, n_rev)


@pytest.mark.gpu
def test_gpu_records_follow_the_documented_rules():
    s, pb = _record_rule_set()
    ctx = helpers.gpu_context(s)
    n_multi, n_fallback, n_rev = check_record_rules(ctx, s, pb, ctx.get_contig_segments())
    assert n_multi > 100 and n_fallback > 10 and n_rev > 500


# ------------------------------------------------------------------------------------------------------------------------
# a3 (which contig segments a read segment is lifted through) evaluated with numpy from the reference's predicate
# (src/read_alignment_scanner.rs:80-103 + IntRange::intersect_range, int_range.rs:56-58: `other.end >= self.start &&
# other.start < self.end`, i.e. a segment that merely TOUCHES the read's low end counts; SURVEY parity trap 1)

def check_pair_enumeration(ctx, s, pb, segs):
    r = helpers.lift_c(ctx, pb.c, allow_panic=True)
    b = pb.c
    g = lambda p, n: np.ctypeslib.as_array(p, (n,))
    ns = b.n_read_segments
    ctg, pos, cb, cl = g(b.rseg_contig, ns), g(b.rseg_pos, ns), g(b.rseg_cigar_begin, ns), g(b.rseg_cigar_len, ns)
    pool = g(b.cigar, int(b.n_cigar))
    want = set()
    n_touching = 0
    for k in range(ns):
        ops = pool[int(cb[k]): int(cb[k]) + int(cl[k])]
        start = int(pos[k])
        end = start + int(sum(int(x) >> 4 for x in ops if (int(x) & 15) in REF_OPS))
        g0, g1 = int(segs.contig_seg_begin[ctg[k]]), int(segs.contig_seg_begin[ctg[k] + 1])
        for q in range(g0, g1):
            s0, s1 = int(segs.seg_seq_order_start[q]), int(segs.seg_seq_order_end[q])
            if end >= s0 and start < s1:
                want.add((k, q - g0))
                n_touching += end == s0
    assert r.n_pairs == len(want)
    lifted = {(int(r.rec_read_segment[k]), int(r.rec_contig_segment[k])) for k in range(r.n_records) if r.rec_status[k] == 1}
    assert lifted <= want and r.n_lifted == len(lifted)
    return len(want), len(lifted), n_touching


def test_oracle_and_device_code_enumerate_the_pairs_of_the_predicate():
    import emul_lib
    s, pb = _record_rule_set()
    for L in (oracle_lib.load(), emul_lib.load()):
        ctx = abi.Context(L, 0, 1)
        ctx.set_reference(helpers.reference_arrays(s))
        ctx.set_contig_records(s.contig_records)
        n_pairs, n_lifted, _ = check_pair_enumeration(ctx, s, pb, ctx.get_contig_segments())
        assert n_pairs > pb.c.n_reads and n_lifted > 0.8 * n_pairs


@pytest.mark.gpu
def test_gpu_enumerates_the_pairs_of_the_predicate():
    s, pb = _record_rule_set()
    ctx = helpers.gpu_context(s)
    n_pairs, n_lifted, _ = check_pair_enumeration(ctx, s, pb, ctx.get_contig_segments())
    assert n_pairs > pb.c.n_reads and n_lifted > 0.8 * n_pairs
