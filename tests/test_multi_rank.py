"""N>1 host logic on CPU: two gloo ranks shard one input set by the reference's (contig x window) units, each lifts only
its units, and the ordered concatenation equals the single-process result.  The compute stand-in on CPU is the oracle
(the product has no CPU path); the GPU-marked variant runs the same flow with the CUDA library."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from portello_b200 import abi, shard, synth

FIELDS = ("rec_status", "rec_tid", "rec_pos", "rec_mapq", "rec_flag", "rec_bin", "rec_need_flip", "cigar")
KW = dict(seed=77, n_chrom=2, chrom_len=600_000, contigs_per_chrom=2, n_reads=2500, junction_per_mb=6.0)
WINDOW = 100_000  # small windows so that the tiny data set has many units


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, use_gpu, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = synth.make("tiny", **KW)
        rr = s.read_records
        n = rr.n_reads
        tid = np.ctypeslib.as_array(rr.tid, (n,))
        pos = np.ctypeslib.as_array(rr.pos, (n,))
        clen = [int(s.contig_records.contig_len[i]) for i in range(s.n_contigs)]
        units = shard.window_units(clen, tid, pos, WINDOW)
        owner = shard.assign(units, world)
        ctx = helpers.gpu_context(s) if use_gpu else helpers.oracle_context(s, threads=1)
        mine = {}
        for ui, u in enumerate(units):
            if owner[ui] != rank or u.n_reads == 0:
                continue
            pb = helpers.pack(s, u.first_read, u.n_reads)
            r = helpers.lift_c(ctx, pb.c)
            mine[ui] = {f: getattr(r, f) for f in FIELDS}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)  # control plane only (results to the host that writes the BAM)
        if rank == 0:
            merged = {}
            for g in gathered:
                merged.update(g)
            order = sorted(merged)
            cat = {f: np.concatenate([merged[u][f] for u in order]) for f in FIELDS}
            whole = helpers.lift_c(ctx, helpers.pack(s).c)
            ok = all(np.array_equal(cat[f], getattr(whole, f)) for f in FIELDS)
            loads = np.bincount(owner, weights=[u.n_reads for u in units], minlength=world)
            with open(out_path, "w") as fh:
                fh.write(f"{int(ok)} {len(units)} {int(loads.min())} {int(loads.max())} {sum(len(g) for g in gathered)}")
    finally:
        dist.destroy_process_group()


def _run(tmp_path, use_gpu):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), use_gpu, out), nprocs=2, join=True)
    ok, n_units, lo, hi, n_done = map(int, open(out).read().split())
    assert ok == 1, "sharded + ordered-gather result differs from the single-process result"
    assert n_units >= 8 and n_done >= 6
    assert hi <= 1.5 * max(lo, 1), (lo, hi)  # LPT keeps the two ranks balanced


def test_two_rank_sharding_matches_single_process_cpu(tmp_path):
    _run(tmp_path, use_gpu=False)


@pytest.mark.gpu
def test_two_rank_sharding_matches_single_process_gpu(tmp_path):
    _run(tmp_path, use_gpu=True)


def _bench_module():
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_tests", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bench_for_tests"] = mod
    spec.loader.exec_module(mod)
    return mod


def _shard_digest_worker(rank, world, port, out_path):
    """bench.py's own N > 1 flow on CPU: plan the whole read set, split it by the reference's work units (LPT), generate and
    lift ONLY this rank's units (the oracle stands in for the GPU), digest the shard with global read indices, gather the
    digests on the control plane: their sum must be the digest of the unsharded run."""
    import argparse
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from portello_b200.digest import Digest
        bench = _bench_module()
        args = argparse.Namespace(workload="tiny", reads=2000)
        s, gri, units, owner, loads = bench.make_shard(args, rank, world, assign=shard.assign)
        assert s.read_records.n_reads == len(gri) == int(loads[rank])
        pb = helpers.pack(s)
        r = helpers.lift_c(helpers.oracle_context(s, threads=1), pb.c)
        d = Digest().add(r, np.ctypeslib.as_array(pb.c.read_seg_begin, (pb.c.n_reads + 1,)), gri)
        parts = [None] * world
        dist.all_gather_object(parts, d.as_tuple())
        if rank == 0:
            tot = Digest()
            for p in parts:
                o = Digest()
                o.acc, o.n_records, o.n_ops, o.n_reads = p
                tot.merge(o)
            s1, gri1, _, _, _ = bench.make_shard(args, 0, 1)
            pb1 = helpers.pack(s1)
            r1 = helpers.lift_c(helpers.oracle_context(s1, threads=1), pb1.c)
            whole = Digest().add(r1, np.ctypeslib.as_array(pb1.c.read_seg_begin, (pb1.c.n_reads + 1,)), gri1)
            with open(out_path, "w") as fh:
                fh.write(f"{int(tot == whole)} {tot.n_reads} {int(loads.min())} {int(loads.max())}")
    finally:
        dist.destroy_process_group()


def test_bench_sharding_digest_is_the_unsharded_digest(tmp_path):
    out = str(tmp_path / "digest.txt")
    mp.spawn(_shard_digest_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ok, n, lo, hi = map(int, open(out).read().split())
    assert ok == 1 and n == 2000
    assert lo + hi == 2000 and hi <= 1.5 * lo
