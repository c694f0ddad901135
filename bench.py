#!/usr/bin/env python
"""bench.py — read alignments lifted per second on B200 (BASELINE.json metric), next to the CPU path.

Workload (default): BASELINE.json configs[2], the whole-genome synthetic diploid assembly (24 x 130 Mb reference, ~1000
contigs, 6,000,000 15 kb HiFi reads).  A "step" is one pass of the liftover hot path over the WHOLE read set.
`--gpus N` is configs[3]: the SAME set split over N ranks by the reference's own work units, (contig x <= 20 Mb window)
(src/read_alignment_scanner.rs:495-535,606-660), bin-packed by read count (ptl_shard_units, greedy LPT); every rank
generates, lifts and checks only its own units, results are ordered by unit = the reference's record order, no
collective on the data path (torch.distributed/NCCL carries the barrier, the timing reductions and the digests only).

  value     device-resident throughput: the rank's shard is already in HBM, K x ptl_lift_run timed with CUDA events on the
            library's stream (pair enumeration -> lift -> finalize -> record emission); whole job = all pairs / max over ranks.
  e2e       the same work through the public C-ABI call a host makes (ptl_lift_submit / ptl_lift_wait) from pinned HOST
            buffers, H2D and D2H inside the timed region, pipelined over 3 slots in chunks of 131072 reads.
  parity    100 % of the records of every rank's shard: an order-sensitive digest (portello_b200/digest.py) of the CUDA
            result equals the digest of the CPU oracle's result; the sum over ranks is printed (identical for every N).
  roofline  lift_pairs kernel: algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json HBM GB/s.
  cpu_baseline  the oracle (CPU restatement of the reference) on a bounded sample, all host threads.

`--impl reference` times the oracle instead (the reference is Rust and cannot be built in this image); it loads the
oracle and the generator only, never libportello_b200.so.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOAD_DOC = {
    "config1": "BASELINE.json configs[0], 1 Mb reference, 2 contigs, 20k 15 kb HiFi reads",
    "chr20": "BASELINE.json configs[1], chr20-scale synthetic, ~40 contig alignments carrying SVs, 15 kb HiFi reads",
    "wg": "BASELINE.json configs[2]/[3], whole-genome synthetic diploid assembly, ~500 contigs per haplotype, 30x 15 kb HiFi reads",
    "stress": "BASELINE.json configs[4], fragmented assembly, 100 kb reads with dense clustered indels and SA segments",
}
METRIC = "read alignments lifted/sec"
UNIT = "alignments/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="wg")
    ap.add_argument("--reads", type=int, default=0, help="override the size of the read set (default: the workload's own)")
    ap.add_argument("--chunk", type=int, default=131072, help="reads per submitted batch in the e2e pipeline")
    ap.add_argument("--pack-chunk", type=int, default=32768, help="reads per packed batch in e2e_with_pack (the packer threads take turns chunk by chunk: "
                    "47 chunks of 131072 reads on 15 threads are 4 rounds for 3.1 rounds of work)")
    ap.add_argument("--slots", type=int, default=6, help="batch slots (CUDA streams) of the e2e pipeline (3 / 4 / 6 slots: 25.4 / 24.5 / 23.7 ms per step)")
    ap.add_argument("--value-slots", type=int, default=0, help="resident sub-batches of the `value` step (0 = by shard size)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-assemble", action="store_true", help="skip the record-assembly measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-digest parity check (profiling runs only)")
    ap.add_argument("--zero-copy", default="auto", choices=["auto", "on", "off"])
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        # ONE poller per job (rank 0 watches every GPU of the job)
        self.idx, self.rows, self.p = ",".join(str(i) for i in gpu_indices), [], None

    def start(self):
        if not self.idx:
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", self.idx],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=2)
            except Exception:
                self.p.kill()

    def summary(self, windows):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if not any(a <= t <= b for a, b in windows) or len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(cnt, n_table_entries, n_segments):
    """sum_p [32 + 4 n_in + s + 16 + 4 n_out] (+ 8 T + 24 C once per run, reported separately)."""
    per_pairs = 48 * cnt["n_pairs"] + 4 * cnt["n_in_ops"] + cnt["base_bytes"] + 4 * cnt["n_out_ops"]
    return per_pairs, 8 * n_table_entries + 24 * n_segments


# ------------------------------------------------------------------------------------------------ the data set and its shards
def workload_config(args, s, n_total, n_units, world):
    return (f"{args.workload}: {WORKLOAD_DOC.get(args.workload, 'custom')} ({s.n_chrom} x {int(s.chrom_len[0])} bp reference, "
            f"{s.contig_records.n_records} contig alignment records, {n_total} reads in {n_units} (contig x <=20 Mb window) units)")


def make_shard(args, rank, world, host_alloc=None, host_free=None, region_segments=None, assign=None):
    """Plan the whole read set, split it by the reference's work units, generate only this rank's reads."""
    from portello_b200 import shard, synth

    kw = {"defer_reads": 1}
    if args.reads:
        kw["n_reads"] = args.reads
    if world > 1:
        kw["n_threads"] = max(1, (os.cpu_count() or 1) // world)  # (every rank builds the reference and the contigs: share the cores)
    s = synth.make(args.workload, host_alloc=host_alloc, host_free=host_free, **kw)
    plan_contig, plan_pos = s.plan()
    units = shard.window_units(s.contig_lengths(), plan_contig, plan_pos, region_segments=region_segments)
    owner = assign(units, world) if assign else np.zeros(len(units), np.uint32)
    ranges, gri = shard.rank_ranges(units, owner, rank)
    merged = []  # adjacent ranges of the BAM order (consecutive units of a contig) as one
    for a, n in ranges:
        if merged and merged[-1][0] + merged[-1][1] == a:
            merged[-1][1] += n
        else:
            merged.append([a, n])
    s.generate_reads([tuple(r) for r in merged])
    loads = np.bincount(np.asarray(owner, np.int64), weights=[u.n_reads for u in units], minlength=world)
    return s, gri, units, owner, loads


def time_oracle(octx, batch_c, reps=1):
    from portello_b200 import abi
    import oracle_lib

    O = oracle_lib.load()
    best, pairs, r = None, 0, None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = O._lift_submit(octx.h, 0, C.byref(batch_c))
        r = abi.ResultC()
        rc2 = O._lift_wait(octx.h, 0, C.byref(r))
        dt = time.perf_counter() - t0
        assert rc == 0 and rc2 == 0, (rc, rc2)
        pairs = int(r.n_pairs)
        best = dt if best is None else min(best, dt)
    return best, pairs, r


def oracle_context_for(s, threads, faithful):
    """Oracle context without touching the product library."""
    import oracle_lib
    from portello_b200 import abi

    O = oracle_lib.load()
    ctx = abi.Context(O, 0, 1)
    ctx.set_reference(s.reference_arrays())
    ctx.set_contig_records(s.contig_records)
    O.dll.ptl_oracle_set_threads(ctx.h, threads)
    O.dll.ptl_oracle_set_faithful_decode(ctx.h, int(faithful))
    return ctx


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference on all host threads (rank 0 only), on a bounded sample of the
    SAME workload: whole work units of the whole-genome set, taken evenly across the genome until the sample is full."""
    if rank != 0:
        return
    import oracle_lib
    from portello_b200 import shard, synth

    O = oracle_lib.load()

    def regions(size, seg):
        b, e = (C.c_uint64 * 64)(), (C.c_uint64 * 64)()
        n = O.dll.ptl_oracle_fn_region_segments(size, seg, b, e, 64)
        return [(int(b[i]), int(e[i])) for i in range(n)]

    kw = {"defer_reads": 1}
    if args.reads:
        kw["n_reads"] = args.reads
    s = synth.make(args.workload, **kw)
    plan_contig, plan_pos = s.plan()
    n_total = len(plan_contig)
    units = shard.window_units(s.contig_lengths(), plan_contig, plan_pos, region_segments=regions)
    n_sample = min(n_total, args.cpu_sample or 200_000)
    live = [u for u in units if u.n_reads > 0]
    stride = max(1, int(len(live) * (sum(u.n_reads for u in live) / max(len(live), 1)) / max(n_sample, 1)))
    picked, got = [], 0
    for u in live[::stride]:
        if got >= n_sample:
            break
        picked.append((u.first_read, u.n_reads))
        got += u.n_reads
    s.generate_reads(picked)
    threads = os.cpu_count() or 1
    octx = oracle_context_for(s, threads, faithful=True)
    pb = oracle_lib.OraclePackedBatch(s.read_records, 0, s.read_records.n_reads, s.contig_names, threads=threads)
    for _ in range(args.warmup):
        time_oracle(octx, pb.c)
    times, pairs = [], 0
    for _ in range(args.steps):
        dt, pairs, _ = time_oracle(octx, pb.c)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = pairs / (ms / 1e3)
    sample = (f"{pb.c.n_reads} reads ({pairs} pairs) per step = {len(picked)} whole work units taken evenly across the {n_total}-read set; "
              "oracle = C++ restatement of portello's path incl. its per-pair read decode (faithful_decode)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_config(args, s, n_total, len(units), world), "reads_per_step": int(pb.c.n_reads), "host_cores": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ the native arm
def value_slots_for(n_reads: int) -> int:
    """Resident sub-batches of the `value` step.  Three at every shard size: measured on rank-sized shards of the whole-genome
    set (tools/value_slots_sweep.sh), 0.75 M reads take 0.790 / 0.768 / 0.736 ms as 1 / 2 / 3 sub-batches and 1.5 M reads
    1.332 / 1.292 / 1.244 ms -- the streams fill each other's kernel tails even when each launch is small."""
    return 3 if n_reads >= 3 else 1


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import helpers
    from portello_b200 import abi, lib, shard
    from portello_b200.digest import Digest

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the liftover path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM) if world > 1 else x

    def all_ranks(x: float):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def gather_objects(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)  # control plane only (digests, counters)
        return out

    L = lib.load(build_if_missing=False)
    threads_here = max(1, (os.cpu_count() or 1) // world)

    # ---- data: ONE read set, split by the reference's work units; this rank generates and lifts only its own units
    t0 = time.time()
    alloc = C.cast(L.dll.ptl_host_alloc, C.c_void_p)
    free = C.cast(L.dll.ptl_host_free, C.c_void_p)
    s, gri, units, owner, loads = make_shard(args, rank, world, alloc, free, assign=shard.assign)  # packed bases in pinned, mapped memory
    t_gen = time.time() - t0
    n_reads = s.read_records.n_reads
    n_total = int(sum(u.n_reads for u in units))

    t0 = time.time()
    ctx = lib.GpuContext(local_rank, n_slots=max(args.slots, 2))
    ctx.set_reference(s.reference_arrays())
    ctx.set_contig_records(s.contig_records)
    segs = ctx.get_contig_segments()
    n_segments = len(segs.seg_pos)
    n_table = sum(len(ctx.get_segment_table(g)[0]) for g in range(n_segments))
    t_setup = time.time() - t0

    t0 = time.time()
    # value: the shard as V resident sub-batches, one per slot, lifted CONCURRENTLY on the slots' streams (every kernel of the
    # path is latency-bound, not throughput-bound: a second and third stream fill the tails; +6 % over one resident batch)
    V = args.value_slots or value_slots_for(n_reads)
    V = max(1, min(V, max(args.slots, 2), n_reads))
    vb = [(k * (n_reads // V), (n_reads // V) if k < V - 1 else n_reads - (V - 1) * (n_reads // V)) for k in range(V)]
    value_parts = [lib.PackedBatch(L, s.read_records, a, c, s.contig_names, pinned=True, windows=segs) for a, c in vb]
    chunk_sets = {
        False: [lib.PackedBatch(L, s.read_records, a, min(args.chunk, n_reads - a), s.contig_names, pinned=True)
                for a in range(0, n_reads, args.chunk)],
        # indel windows (ptl_pack_batch_ex) for the read segments that pair with a reverse-strand contig segment
        True: [lib.PackedBatch(L, s.read_records, a, min(args.chunk, n_reads - a), s.contig_names, pinned=True, windows=segs)
               for a in range(0, n_reads, args.chunk)],
    }
    t_pack = time.time() - t0
    assert sum(p.c.n_reads for p in value_parts) == n_reads, "the generator emits primary records only"

    sampler = ClockSampler(range(world) if rank == 0 else [])
    sampler.start()
    windows = []

    # ---------------------------------------------------------------- value: device-resident
    dev = torch.device("cuda", local_rank)
    streams = [torch.cuda.ExternalStream(ctx.stream(k), device=dev) for k in range(V)]
    for k in range(V):
        ctx.upload(value_parts[k].c, k)

    def value_step():
        # fork: every slot's stream starts behind stream 0's position; join: stream 0 waits for all of them
        ev = torch.cuda.Event()
        ev.record(streams[0])
        for k in range(1, V):
            streams[k].wait_event(ev)
        for k in range(V):
            ctx.run(k)
        for k in range(1, V):
            e = torch.cuda.Event()
            e.record(streams[k])
            streams[0].wait_event(e)

    for _ in range(max(args.warmup, 3)):
        value_step()
    for k in range(V):
        ctx.counters(k)  # syncs; also settles buffer capacities (a capacity re-run can only happen here)
    value_step()
    for k in range(V):
        ctx.counters(k)
    barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    ev0.record(streams[0])
    for _ in range(args.steps):
        value_step()
    ev1.record(streams[0])
    ev1.synchronize()
    barrier()
    w1 = time.time()
    windows.append((w0, w1))
    launches_value = ctx.launch_count() - launches0
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    cnt = {}
    for k in range(V):
        for key, v in ctx.counters(k).items():
            cnt[key] = cnt.get(key, 0) + int(v)
    # per-stage times: the V launches of a step one after the other (not concurrently), each bracketed by its own events on
    # its stream, averaged over a few extra (untimed for `value`) passes; stage_ms = the sum over the V launches
    stage_acc = {}
    for _ in range(min(args.steps, 5)):
        tot = {}
        for k in range(V):
            ctx.run(k)
            for key, v in ctx.kernel_times(k).items():
                tot[key] = tot.get(key, 0.0) + v
        for key, v in tot.items():
            stage_acc.setdefault(key, []).append(v)
    stage_ms = {k: float(np.mean(v)) for k, v in stage_acc.items()}
    dev_ms_max = max_over_ranks(dev_ms)
    rank_ms = all_ranks(dev_ms)
    rank_pairs = all_ranks(float(cnt["n_pairs"]))
    pairs_total = sum(rank_pairs)
    value = pairs_total / (dev_ms_max / 1e3)

    # ---------------------------------------------------------------- parity: 100 % of this rank's records vs the oracle
    # (outside every timed region; the oracle is the checker of the measured path, never the thing measured)
    def batch_seg_begin(bc):
        return np.ctypeslib.as_array(bc.read_seg_begin, (bc.n_reads + 1,))

    parity, digest_total, oracle_s = "skipped (--no-parity)", None, None
    d_oracle = None
    if not args.no_parity:
        d_gpu = Digest()
        for k in range(V):
            rg = ctx.download(k, copy=False)
            d_gpu.add(rg, batch_seg_begin(value_parts[k].c), gri[vb[k][0]: vb[k][0] + vb[k][1]])
        octx = helpers.oracle_context(s, threads=threads_here)
        d_oracle = Digest()
        oracle_s = 0.0
        for k in range(V):
            t0 = time.perf_counter()
            _, _, ro_c = time_oracle(octx, value_parts[k].c)
            oracle_s += time.perf_counter() - t0
            d_oracle.add(abi.Result.from_c(ro_c, copy=False), batch_seg_begin(value_parts[k].c), gri[vb[k][0]: vb[k][0] + vb[k][1]])
        del octx
        if d_gpu != d_oracle:
            raise SystemExit(f"PARITY FAILURE vs oracle on rank {rank}: CUDA {d_gpu.hex()} != oracle {d_oracle.hex()}")
        parts = gather_objects(d_gpu.as_tuple())
        tot = Digest()
        for p in parts:
            o = Digest()
            o.acc, o.n_records, o.n_ops, o.n_reads = p
            tot.merge(o)
        digest_total = tot.hex()
        parity = (f"100 % of {tot.n_reads} reads ({tot.n_records} records): order-sensitive digest of the CUDA result == the CPU oracle's on every rank"
                  + (f"; digest of the {world} shards in unit order = {digest_total}" if world > 1 else ""))

    # ---------------------------------------------------------------- e2e: host buffers -> records on the host
    def e2e_step(chunks, on_result=None):
        n_slots = max(args.slots, 2)
        inflight = [None] * n_slots
        recs = 0

        def drain(sl):
            nonlocal recs
            r = ctx.wait_c(sl)
            recs += r.n_records
            if on_result:
                on_result(inflight[sl], r)
            inflight[sl] = None

        for i, ch in enumerate(chunks):
            sl = i % n_slots
            if inflight[sl] is not None:
                drain(sl)
            ctx.submit_c(ch[1].c, sl)
            inflight[sl] = ch
        for k in range(n_slots):
            sl = (len(chunks) + k) % n_slots
            if inflight[sl] is not None:
                drain(sl)
        return recs

    # base-transfer modes of the e2e call: (zero-copy packed bases, indel windows travelling with the batch)
    MODES = {"seq_bulk_upload": (False, False), "seq_zero_copy": (True, False), "seq_zero_copy_indel_windows": (True, True)}

    def chunks_of(mode):
        return [(a * args.chunk, ch) for a, ch in enumerate(chunk_sets[MODES[mode][1]])]

    def measure_e2e(mode):
        zero_copy, _ = MODES[mode]
        chunks = chunks_of(mode)
        ctx.set_seq_zero_copy(zero_copy)
        steps = args.steps if zero_copy else max(1, min(args.steps, 3))  # (bulk upload moves every packed base: seconds per step)
        for _ in range(max(args.warmup, 3) if zero_copy else 1):
            e2e_step(chunks)
        barrier()
        l0 = ctx.launch_count()
        t0 = time.time()
        p0 = time.perf_counter()
        for _ in range(steps):
            n_rec = e2e_step(chunks)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - p0) / steps
        barrier()
        windows.append((t0, time.time()))
        return dt, n_rec, (ctx.launch_count() - l0) // steps, steps

    def h2d_bytes(mode):
        zero_copy, use_win = MODES[mode]
        chunks = chunk_sets[use_win]
        small = sum(int(ch.c.n_reads) * (2 + 1 + 2 + 4 + 8 + 4) + int(ch.c.n_read_segments) * (4 + 8 + 1 + 8 + 4) + int(ch.c.n_cigar) * 4 for ch in chunks)
        win = sum((int(ch.c.n_read_segments) + 1) * 4 + int(ch.c.n_indel_win) * 8 for ch in chunks) if use_win else 0
        seq = 0 if zero_copy else sum(int(ch.c.seq4_bytes) for ch in chunks)
        return small + win + seq

    def e2e_digest(mode):
        """The e2e call path itself, chunk by chunk, digested (outside the timed region)."""
        ctx.set_seq_zero_copy(MODES[mode][0])
        d = Digest()

        def on_result(ch, r):
            a, pbk = ch
            d.add(abi.Result.from_c(r, copy=False), batch_seg_begin(pbk.c), gri[a:a + pbk.c.n_reads])

        e2e_step(chunks_of(mode), on_result)
        return d

    e2e_runs = {}
    modes = {"auto": list(MODES), "on": ["seq_zero_copy", "seq_zero_copy_indel_windows"], "off": ["seq_bulk_upload"]}[args.zero_copy]
    if not args.no_parity:
        for m in modes:
            if m == "seq_bulk_upload" and n_reads > 2_000_000:
                continue  # (45 GB of H2D for a digest; the mode is covered by the GPU test-suite)
            dm = e2e_digest(m)
            if dm != d_oracle:
                raise SystemExit(f"PARITY FAILURE: e2e mode {m} on rank {rank}: {dm.hex()} != oracle {d_oracle.hex()}")
    for m in modes:
        dt, n_rec, nl, st = measure_e2e(m)
        e2e_runs[m] = (max_over_ranks(dt), n_rec, nl, st, dt)
    best_mode = min(e2e_runs, key=lambda k: e2e_runs[k][0])
    e2e_dt, n_rec, launches_e2e, _, e2e_dt_here = e2e_runs[best_mode]
    d2h = int(n_rec) * (1 + 4 + 4 + 4 + 8 + 1 + 2 + 2 + 1 + 8) + int(cnt["n_out_ops"]) * 4 + (n_reads + 1) * 4
    e2e_value = pairs_total / e2e_dt
    rank_e2e_ms = all_ranks(e2e_dt_here * 1e3)
    h2d_here = h2d_bytes(best_mode)
    h2d_total, d2h_total = sum_over_ranks(float(h2d_here)), sum_over_ranks(float(d2h))

    def measure_copy_only(nb_h2d, nb_d2h, reps=5):
        """The same bytes as one e2e step, as two plain pinned copies (one each way, concurrently) on every rank at once and
        nothing else: the box's DMA floor for the step.  e2e / this = how much of the step is not PCIe."""
        hb_in, hb_out = torch.empty(nb_h2d, dtype=torch.uint8).pin_memory(), torch.empty(nb_d2h, dtype=torch.uint8).pin_memory()
        db_in, db_out = torch.empty(nb_h2d, dtype=torch.uint8, device="cuda"), torch.zeros(nb_d2h, dtype=torch.uint8, device="cuda")
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def once():
            with torch.cuda.stream(s_in):
                db_in.copy_(hb_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                hb_out.copy_(db_out, non_blocking=True)

        once()
        torch.cuda.synchronize()
        barrier()
        p0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - p0) / reps
        barrier()
        return dt

    copy_dt_here = measure_copy_only(int(h2d_here), int(d2h))
    copy_dt = max_over_ranks(copy_dt_here)
    # ---------------------------------------------------------------- e2e_with_pack: the host packer on the clock
    # decoded records -> ptl_pack_batch_into (split segments, SA parse, indel windows) on P host threads, into a ring of
    # reusable pinned batches, overlapped with submit / wait of earlier chunks (one packer thread per slot).
    def measure_e2e_with_pack(n_pack):
        import queue
        ctx.set_seq_zero_copy(True)
        n_slots = max(args.slots, 2)
        pc = max(1, args.pack_chunk)
        bounds = [(a, min(pc, n_reads - a)) for a in range(0, n_reads, pc)]
        ring = [lib.PackedBatch(L, s.read_records, 0, min(pc, n_reads), s.contig_names, pinned=True, windows=segs) for _ in range(n_slots + n_pack + 1)]
        pack_s = [0.0] * n_pack

        def one_pass():
            # chunk i is packed into ring[i % R], once the consumer has released chunk i - R: a packer that runs ahead waits
            # instead of taking the buffer a chunk in front of it needs (free-list hand-out could starve the next chunk)
            R = len(ring)
            cv = threading.Condition()
            released = [0]  # chunks 0 .. released-1 are back
            ready = [queue.Queue(1) for _ in bounds]

            def packer(t):
                for i in range(t, len(bounds), n_pack):
                    with cv:
                        cv.wait_for(lambda: i - R < released[0])
                    t0 = time.perf_counter()
                    ring[i % R].repack(bounds[i][0], bounds[i][1])
                    pack_s[t] += time.perf_counter() - t0
                    ready[i].put(i % R)

            def release():
                with cv:
                    released[0] += 1
                    cv.notify_all()

            th = [threading.Thread(target=packer, args=(t,)) for t in range(n_pack)]
            [t.start() for t in th]
            inflight = [None] * n_slots
            recs = 0
            for i in range(len(bounds)):
                b = ready[i].get()
                sl = i % n_slots
                if inflight[sl] is not None:
                    recs += ctx.wait_c(sl).n_records
                    release()
                ctx.submit_c(ring[b].c, sl)
                inflight[sl] = b
            for k in range(n_slots):
                sl = (len(bounds) + k) % n_slots
                if inflight[sl] is not None:
                    recs += ctx.wait_c(sl).n_records
                    inflight[sl] = None
                    release()
            [t.join() for t in th]
            return recs

        one_pass()
        for t in range(n_pack):
            pack_s[t] = 0.0
        steps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.time()
        p0 = time.perf_counter()
        for _ in range(steps):
            n_rec2 = one_pass()
        dt = (time.perf_counter() - p0) / steps
        barrier()
        windows.append((t0, time.time()))
        assert n_rec2 == n_rec
        return dt, n_reads * steps / max(sum(pack_s), 1e-9)

    n_pack = max(1, threads_here - 1)
    pk_dt, pk_rate_thread = measure_e2e_with_pack(n_pack)
    pk_dt_max = max_over_ranks(pk_dt)
    e2e_with_pack = {"value": pairs_total / pk_dt_max, "unit": UNIT, "ms_per_step": pk_dt_max * 1e3, "packer_threads_per_rank": n_pack, "reads_per_packed_batch": args.pack_chunk,
                     "packer_reads_per_s_per_thread": pk_rate_thread, "ratio_to_e2e": (pairs_total / pk_dt_max) / e2e_value,
                     "what": "as e2e, plus the host packer inside the timed region: ptl_pack_batch_into (split segments, SA parse, indel windows) on "
                             "host threads into a ring of reusable pinned batches, overlapped with the GPU; a real run decodes BAM before this, "
                             "at ~1e4 reads/s per inflate thread (tools/cli_e2e.py measures that path)"}
    sampler.stop()

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6.65 TB/s"
    alg_bytes, table_bytes = algorithmic_bytes(cnt, n_table, n_segments)
    lift_ms_sum = stage_ms.get("lift_pairs", float("nan"))  # over the V launches of a step
    lift_ms_avg = lift_ms_sum / V
    achieved = (alg_bytes / V) / (lift_ms_avg / 1e3) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel on the same workload, from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["lift_pairs_kernel"]
        if tj["workload"] == args.workload and tj["reads_per_launch"] == vb[0][1]:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "lift_pairs_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes // V), "table_bytes_once": int(table_bytes),
                "kernel_ms": lift_ms_avg, "launches_per_step": V, "reads_per_launch": int(vb[0][1]),
                "note": "kernel_ms = CUDA events around the lift stage (lift_pairs_kernel + its two small worklist kernels) of one launch, the V launches "
                        "of a step run one after the other for this measurement; `value` runs them concurrently on V streams",
                "stage_ms": stage_ms, "whole_step_frac": alg_bytes / (dev_ms / 1e3) / 1e9 / peak}

    extra = {}
    if rank == 0 and world == 1 and not args.no_assemble:
        import bench_records
        try:  # (a by-product of the run: the record path must not cost the liftover line if it fails)
            extra = bench_records.measure(args, ctx, s, L, chunk_sets[False][0], peak)
        except Exception as e:  # noqa: BLE001
            extra = {"assemble_records": {"error": f"{type(e).__name__}: {e}"}}
            print(f"bench_records failed: {e}", file=sys.stderr)

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_sample = args.cpu_sample or min(n_reads, max(20_000, 2500 * threads))
        sub = lib.PackedBatch(L, s.read_records, 0, n_sample, s.contig_names)
        octx = helpers.oracle_context(s, threads=threads, faithful=True)
        time_oracle(octx, sub.c)
        dt, pairs, _ = time_oracle(octx, sub.c, reps=3)
        octx1 = helpers.oracle_context(s, threads=threads, faithful=False)
        dt_lean, _, _ = time_oracle(octx1, sub.c, reps=3)
        cpu = {"value": pairs / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n_sample} reads ({pairs} pairs) of this workload, best of 3; oracle = C++ restatement incl. the reference's "
                         f"per-pair read decode; without that decode (lazy base access): {pairs / dt_lean:.3e} {UNIT}"
                         + (f"; the full-set parity run (lean, {threads_here} threads) took {oracle_s:.1f} s for {n_reads} reads" if oracle_s else "")}

    if rank == 0:
        mean_ms = float(np.mean(rank_ms))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_config(args, s, n_total, len(units), world),
                       "reads_total": n_total, "reads_per_rank": [int(x) for x in loads], "pairs_per_rank": [int(x) for x in rank_pairs],
                       "pairs_per_step": int(pairs_total),
                       "value_pipeline": f"{V} resident sub-batches on {V} slots, lifted concurrently on their streams",
                       "l2_policy": "inputs larger than L2 (CIGAR pools + op scratch > 126 MB per step)",
                       "sharding": ("one read set split over the ranks by (contig x <=20 Mb window) work units, greedy LPT on read counts "
                                    "(ptl_shard_units); results ordered by unit; no collective on the data path") if world > 1 else "single GPU: all units",
                       "rank_ms_per_step": [round(v, 4) for v in rank_ms],
                       "imbalance_max_over_mean": {"device_ms": round(max(rank_ms) / mean_ms, 4), "reads": round(float(loads.max() / max(loads.mean(), 1)), 4),
                                                   "pairs": round(max(rank_pairs) / max(np.mean(rank_pairs), 1), 4)},
                       "rank_e2e_ms_per_step": [round(v, 3) for v in rank_e2e_ms],
                       "e2e_pipeline": f"batches of {args.chunk} reads over {max(args.slots, 2)} slots per rank", "e2e_base_transfer": best_mode,
                       "parity": parity, "digest": digest_total,
                       "host": {"cores": os.cpu_count(), "threads_per_rank": threads_here},
                       "setup_s": {"generate": round(t_gen, 1), "reference_and_tables": round(t_setup, 1), "pack": round(t_pack, 1)}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_total), "d2h_bytes_per_step": int(d2h_total),
                    "ms_per_step": e2e_dt * 1e3,
                    "pcie_gbs_per_rank": {"h2d": round(h2d_here / e2e_dt_here / 1e9, 2), "d2h": round(d2h / e2e_dt_here / 1e9, 2)},
                    "copy_only": {"ms_per_step": copy_dt * 1e3, "gbs_per_rank": {"h2d": round(h2d_here / copy_dt_here / 1e9, 2), "d2h": round(d2h / copy_dt_here / 1e9, 2)},
                                  "aggregate_gbs": round((h2d_total + d2h_total) / copy_dt / 1e9, 1), "frac_of_e2e": copy_dt / e2e_dt,
                                  "what": "the step's H2D and D2H bytes as two plain pinned copies per rank, all ranks at once, nothing else: "
                                          "the DMA floor of the e2e step on this box"},
                    "alternatives": {k: {"value": pairs_total / v[0], "ms_per_step": v[0] * 1e3, "h2d_bytes_per_step_rank0": int(h2d_bytes(k)), "steps": v[3]}
                                     for k, v in e2e_runs.items()}},
            "e2e_with_pack": e2e_with_pack,
            "gpu_launches": int(launches_value),
            "gpu_launches_e2e_per_step": int(launches_e2e),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(windows),
            "counters": cnt,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
