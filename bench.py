#!/usr/bin/env python
"""bench.py — read alignments lifted per second on B200 (BASELINE.json metric), next to the CPU path.

A "step" is one pass of the liftover hot path over one synthetic batch: BASELINE.json configs[1]
(chr20-scale: 64 Mb reference, ~40 contig->ref alignments carrying SVs, 1,000,000 HiFi reads) per GPU.

  value     device-resident throughput: the batch is already in HBM, K x ptl_lift_run timed with CUDA events on the
            library's stream (pair enumeration -> lift -> finalize -> record emission), max over ranks.
  e2e       the same work through the public C-ABI call a host makes (ptl_lift_submit / ptl_lift_wait) from pinned HOST
            buffers, H2D and D2H inside the timed region, pipelined over 3 slots.
  roofline  lift_pairs kernel: algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json HBM GB/s.
  cpu_baseline  the oracle (CPU restatement of the reference) on a bounded sample, all host threads.

`--impl reference` times the oracle instead (the reference is Rust and cannot be built in this image).
Multi-GPU: one process per GPU (torchrun), weak scaling, reads shard by contig set with no collective on the data path;
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks reduction of the timings.
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

import numpy as np  # noqa: E402

WORKLOAD_DOC = {
    "config1": "BASELINE.json configs[0], 1 Mb reference, 2 contigs, 20k 15 kb HiFi reads",
    "chr20": "BASELINE.json configs[1], chr20-scale synthetic, ~40 contig alignments carrying SVs, 15 kb HiFi reads",
    "wg": "BASELINE.json configs[2], whole-genome synthetic diploid assembly, ~500 contigs per haplotype, 15 kb HiFi reads",
    "stress": "BASELINE.json configs[4], fragmented assembly, 100 kb reads with dense clustered indels and SA segments",
}
METRIC = "read alignments lifted/sec"
UNIT = "alignments/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="chr20")
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (default: the workload's own)")
    ap.add_argument("--chunk", type=int, default=131072, help="reads per submitted batch in the e2e pipeline")
    ap.add_argument("--slots", type=int, default=3, help="batch slots (CUDA streams) of the e2e pipeline")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-assemble", action="store_true", help="skip the record-assembly (bases) measurement")
    ap.add_argument("--zero-copy", default="auto", choices=["auto", "on", "off"])
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        # ONE poller per job (rank 0 watches every GPU of the job): a poller per rank made 8 nvidia-smi processes query the
        # driver every 100 ms, and at N = 8 some rank's 16 ms timed region caught a multi-millisecond stall almost every run
        # (max over ranks 1.07 ms per step against 0.79 ms on every rank's own stage events).
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


def oracle_ctx_for(s, threads, faithful=True):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers

    return helpers.oracle_context(s, threads=threads, faithful=faithful)


def time_oracle(octx, batch_c, reps=1):
    from portello_b200 import abi
    import oracle_lib

    O = oracle_lib.load()
    best, pairs = None, 0
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = O._lift_submit(octx.h, 0, C.byref(batch_c))
        r = abi.ResultC()
        rc2 = O._lift_wait(octx.h, 0, C.byref(r))
        dt = time.perf_counter() - t0
        assert rc == 0 and rc2 == 0, (rc, rc2)
        pairs = int(r.n_pairs)
        best = dt if best is None else min(best, dt)
    return best, pairs


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference on all host threads (rank 0 only)."""
    if rank != 0:
        return
    from portello_b200 import lib, synth

    kw = {}
    n_sample = args.cpu_sample or 100_000
    kw["n_reads"] = min(args.reads or synth.WORKLOADS[args.workload]["n_reads"], n_sample)
    s = synth.make(args.workload, **kw)
    threads = os.cpu_count() or 1
    octx = oracle_ctx_for(s, threads, faithful=True)
    pb = lib.PackedBatch(lib.load(), s.read_records, 0, s.read_records.n_reads, s.contig_names)
    for _ in range(args.warmup):
        time_oracle(octx, pb.c)
    times, pairs = [], 0
    for _ in range(args.steps):
        dt, pairs = time_oracle(octx, pb.c)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = pairs / (ms / 1e3)
    sample = f"{pb.c.n_reads} reads ({pairs} pairs) of the {args.workload} workload per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"{args.workload} (BASELINE.json configs[1] shape), bounded sample", "reads_per_step": int(pb.c.n_reads)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + "; oracle = C++ restatement of portello's path incl. its per-pair read decode (faithful_decode)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


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

    from portello_b200 import abi, lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the liftover path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(x: float):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    L = lib.load(build_if_missing=False)
    # ---- data: weak scaling, no exchange: every rank lifts its own shard of the same shape.  The shards are generated from
    # the SAME seed: with seed + rank the ranks drew different inversion layouts (reverse-strand pairs cost ~3x a forward
    # pair), and the max over ranks measured that synthetic imbalance (N = 8: 1.07 ms on the slowest rank, 0.79 ms on
    # rank 0) instead of the machine.  Per-rank step times are reported in config.rank_ms_per_step.
    kw = {"seed": synth.WORKLOADS[args.workload]["seed"]}
    if args.reads:
        kw["n_reads"] = args.reads
    t0 = time.time()
    alloc = C.cast(L.dll.ptl_host_alloc, C.c_void_p)
    free = C.cast(L.dll.ptl_host_free, C.c_void_p)
    s = synth.make(args.workload, host_alloc=alloc, host_free=free, **kw)  # packed bases generated into pinned, mapped memory
    t_gen = time.time() - t0
    n_reads = s.read_records.n_reads

    ctx = lib.GpuContext(local_rank, n_slots=max(args.slots, 2))
    ctx.set_reference(s.reference_arrays())
    ctx.set_contig_records(s.contig_records)
    segs = ctx.get_contig_segments()
    n_segments = len(segs.seg_pos)
    n_table = sum(len(ctx.get_segment_table(g)[0]) for g in range(n_segments))

    whole = lib.PackedBatch(L, s.read_records, 0, n_reads, s.contig_names, pinned=True)
    win_segs = ctx.get_contig_segments()
    chunk_sets = {
        False: [lib.PackedBatch(L, s.read_records, a, min(args.chunk, n_reads - a), s.contig_names, pinned=True)
                for a in range(0, n_reads, args.chunk)],
        # indel windows (ptl_pack_batch_ex) for the read segments that pair with a reverse-strand contig segment
        True: [lib.PackedBatch(L, s.read_records, a, min(args.chunk, n_reads - a), s.contig_names, pinned=True, windows=win_segs)
               for a in range(0, n_reads, args.chunk)],
    }
    whole_win = lib.PackedBatch(L, s.read_records, 0, n_reads, s.contig_names, pinned=True, windows=win_segs)

    sampler = ClockSampler(range(world) if rank == 0 else [])
    sampler.start()
    windows = []

    # ---------------------------------------------------------------- value: device-resident
    stream = torch.cuda.ExternalStream(ctx.stream(0), device=torch.device("cuda", local_rank))
    ctx.upload(whole.c, 0)
    for _ in range(max(args.warmup, 3)):
        ctx.run(0)
    cnt = ctx.counters(0)  # syncs; also settles buffer capacities (a capacity re-run can only happen here)
    ctx.run(0)
    ctx.counters(0)
    barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    ev0.record(stream)
    lift_ms = []
    for _ in range(args.steps):
        ctx.run(0)
    ev1.record(stream)
    ev1.synchronize()
    barrier()
    w1 = time.time()
    windows.append((w0, w1))
    launches_value = ctx.launch_count() - launches0
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    kt = ctx.kernel_times(0)  # CUDA events around the stages of the LAST timed run, on the library's stream
    cnt = ctx.counters(0)
    # per-stage times averaged over a few extra (untimed for `value`) runs, each bracketed by its own events
    stage_acc = {}
    for _ in range(min(args.steps, 5)):
        ctx.run(0)
        for k, v in ctx.kernel_times(0).items():
            stage_acc.setdefault(k, []).append(v)
    stage_ms = {k: float(np.mean(v)) for k, v in stage_acc.items()}
    dev_ms_max = max_over_ranks(dev_ms)
    rank_ms = all_ranks(dev_ms)
    pairs_total = sum_over_ranks(float(cnt["n_pairs"]))
    value = pairs_total / (dev_ms_max / 1e3)

    # ---------------------------------------------------------------- parity spot check (outside every timed region)
    parity = "skipped"
    if rank == 0:
        # the oracle is used here only as the checker of the measured path (never as the thing measured)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers

        sub = lib.PackedBatch(L, s.read_records, n_reads // 3, min(4000, n_reads - n_reads // 3), s.contig_names)
        rg = helpers.lift_c(ctx, sub.c, slot=1)
        ro = helpers.lift_c(helpers.oracle_context(s, threads=min(8, os.cpu_count() or 1)), sub.c)
        d = rg.diff(ro)
        if d is not None:
            raise SystemExit(f"PARITY FAILURE vs oracle on the bench workload: {d}")
        parity = f"bit-exact vs oracle on {sub.c.n_reads} reads of this workload"

    # ---------------------------------------------------------------- e2e: host buffers -> records on the host
    def e2e_step(chunks):
        n_slots = max(args.slots, 2)
        inflight = [None] * n_slots
        recs = 0
        for i, ch in enumerate(chunks):
            sl = i % n_slots
            if inflight[sl] is not None:
                recs += ctx.wait_c(sl).n_records
            ctx.submit_c(ch.c, sl)
            inflight[sl] = ch
        for k in range(n_slots):
            sl = (len(chunks) + k) % n_slots
            if inflight[sl] is not None:
                recs += ctx.wait_c(sl).n_records
                inflight[sl] = None
        return recs

    def measure_e2e(mode):
        zero_copy, use_win = MODES[mode]
        chunks = chunk_sets[use_win]
        ctx.set_seq_zero_copy(zero_copy)
        for _ in range(max(args.warmup, 3)):
            e2e_step(chunks)
        barrier()
        l0 = ctx.launch_count()
        t0 = time.time()
        p0 = time.perf_counter()
        for _ in range(args.steps):
            n_rec = e2e_step(chunks)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - p0) / args.steps
        barrier()
        windows.append((t0, time.time()))
        return dt, n_rec, ctx.launch_count() - l0

    # base-transfer modes of the e2e call: (zero-copy packed bases, indel windows travelling with the batch)
    MODES = {"seq_bulk_upload": (False, False), "seq_zero_copy": (True, False), "seq_zero_copy_indel_windows": (True, True)}

    def h2d_bytes(mode):
        zero_copy, use_win = MODES[mode]
        chunks = chunk_sets[use_win]
        small = sum(int(ch.c.n_reads) * (2 + 1 + 2 + 4 + 8 + 4) + int(ch.c.n_read_segments) * (4 + 8 + 1 + 8 + 4) + int(ch.c.n_cigar) * 4 for ch in chunks)
        win = sum((int(ch.c.n_read_segments) + 1) * 4 + int(ch.c.n_indel_win) * 8 for ch in chunks) if use_win else 0
        seq = 0 if zero_copy else sum(int(ch.c.seq4_bytes) for ch in chunks)
        return small + win + seq

    def result_digest(mode):
        """Order-sensitive checksum of every result array of the whole batch (outside the timed region)."""
        zero_copy, use_win = MODES[mode]
        ctx.set_seq_zero_copy(zero_copy)
        ctx.submit_c((whole_win if use_win else whole).c, 0)
        r = abi.Result.from_c(ctx.wait_c(0), copy=False)
        acc = np.uint64(1469598103934665603)
        for f in abi.Result.FIELDS:
            a = np.ascontiguousarray(getattr(r, f)).view(np.uint8)
            pad = (-len(a)) % 8
            w = np.concatenate([a, np.zeros(pad, np.uint8)]).view(np.uint64) if pad else a.view(np.uint64)
            acc = np.uint64((int(acc) * 1099511628211 + int(np.bitwise_xor.reduce(w * (np.arange(len(w), dtype=np.uint64) | np.uint64(1))))) & 0xFFFFFFFFFFFFFFFF)
        return int(acc)

    e2e_runs = {}
    modes = {"auto": list(MODES), "on": ["seq_zero_copy", "seq_zero_copy_indel_windows"], "off": ["seq_bulk_upload"]}[args.zero_copy]
    digests = {m: result_digest(m) for m in modes}
    if len(set(digests.values())) != 1:
        raise SystemExit(f"PARITY FAILURE: the base-transfer modes of the full batch differ: {digests}")
    for m in modes:
        dt, n_rec, nl = measure_e2e(m)
        e2e_runs[m] = (max_over_ranks(dt), n_rec, nl)
    best_mode = min(e2e_runs, key=lambda k: e2e_runs[k][0])
    e2e_dt, n_rec, launches_e2e = e2e_runs[best_mode]
    chunks = chunk_sets[False]
    d2h = int(n_rec) * (1 + 4 + 4 + 4 + 8 + 1 + 2 + 2 + 1 + 8) + int(cnt["n_out_ops"]) * 4 + (n_reads + 1) * 4
    e2e_value = pairs_total / e2e_dt
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
    lift_ms_avg = stage_ms.get("lift_pairs", float("nan"))
    achieved = alg_bytes / (lift_ms_avg / 1e3) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel on the same workload, from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["lift_pairs_kernel"]
        if tj["workload"] == args.workload and tj["reads_per_gpu"] == n_reads:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "lift_pairs_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes), "table_bytes_once": int(table_bytes),
                "kernel_ms": lift_ms_avg, "stage_ms": stage_ms,
                "whole_step_frac": alg_bytes / (dev_ms / 1e3) / 1e9 / peak}

    # ---------------------------------------------------------------- next row: record assembly, bases
    # Outside every timed region of the headline metric.  One chunk of the workload, bases + qualities resident in HBM,
    # ptl_assemble_bases in its timing-only mode (no H2D, no D2H): seq revcomp + qual reverse of every output record.
    assemble = None
    if rank == 0 and world == 1 and not args.no_assemble:
        ch = chunk_sets[False][0]
        ctx.set_seq_zero_copy(False)  # the chunk's packed bases are uploaded: the kernel streams them from HBM
        ctx.submit_c(ch.c, 0)
        res = abi.Result.from_c(ctx.wait_c(0), copy=False)
        seq_len = np.ctypeslib.as_array(ch.c.read_seq_len, (ch.c.n_reads,)).astype(np.int64)
        qoff = np.zeros(ch.c.n_reads, np.uint64)
        qoff[1:] = np.cumsum(seq_len[:-1])
        tile = np.random.default_rng(1).integers(0, 94, 1 << 26, dtype=np.uint8)
        qual = np.resize(tile, int(seq_len.sum()))
        o0, _ = ctx.assemble_bases(qual, qoff, 0, flags=abi.ASM_NO_DOWNLOAD)  # uploads the qualities once
        ms = []
        for _ in range(max(args.warmup, 3) + args.steps):
            o, _ = ctx.assemble_bases(None, None, 0, flags=abi.ASM_RESIDENT_QUAL | abi.ASM_NO_DOWNLOAD)
            ms.append(float(o.kernel_ms))
        ms = ms[max(args.warmup, 3):]
        a_ms = float(np.mean(ms))
        a_bytes = int(o.bytes_read + o.bytes_written)
        # parity of this very call path on a slice of the chunk, against the oracle (checker only)
        sub = lib.PackedBatch(L, s.read_records, 0, min(2000, n_reads), s.contig_names)
        sl_len = np.ctypeslib.as_array(sub.c.read_seq_len, (sub.c.n_reads,)).astype(np.int64)
        so = np.zeros(sub.c.n_reads, np.uint64)
        so[1:] = np.cumsum(sl_len[:-1])
        sq = np.resize(tile, int(sl_len.sum()))
        octx_a = helpers.oracle_context(s, threads=min(8, os.cpu_count() or 1))
        helpers.lift_c(octx_a, sub.c)
        helpers.lift_c(ctx, sub.c, slot=1)
        t0 = time.perf_counter()
        _, po = octx_a.assemble_bases(sq, so)
        t_cpu = time.perf_counter() - t0
        _, pg = ctx.assemble_bases(sq, so, 1)
        if not all(np.array_equal(a, b) for a, b in zip(po, pg)):
            raise SystemExit("PARITY FAILURE vs oracle in ptl_assemble_bases on the bench workload")
        a_traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["assemble_records_kernel"]
            if tj["workload"] == args.workload and tj["reads"] == int(ch.c.n_reads):
                a_traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        assemble = {"kernel": "assemble_records_kernel", "what": "seq revcomp (4-bit, no decode) + qual reverse / copy of every output record "
                    "(reverse_alignment_seq_and_qual, src/read_alignment_scanner.rs:125-133); bases + qualities resident in HBM",
                    "records": int(o.n_records), "reads": int(ch.c.n_reads), "flipped_records": int(np.count_nonzero(res.rec_need_flip)),
                    "kernel_ms": a_ms, "records_per_s": o.n_records / (a_ms / 1e3),
                    "roofline": {"bound": "hbm", "achieved": a_bytes / (a_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": a_bytes / (a_ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": a_bytes, "traffic": a_traffic},
                    "cpu_baseline": {"value": len(po[0]) / t_cpu, "unit": "records/s", "cores": 1, "kind": "port",
                                     "sample": f"{sub.c.n_reads} reads of this workload: decode -> rev_comp_in_place -> re-encode as the reference does, "
                                               "incl. the python-side copy of the result"},
                    "parity": f"byte-exact vs oracle on {sub.c.n_reads} reads of this workload"}

    # ---------------------------------------------------------------- the same row, whole BAM records: ptl_assemble_records
    # Every output record as bam_write1 bytes (clone_record tag stripping, field updates, PS/ZM/SA tags, flipped bases and
    # qualities).  Same chunk, everything resident in HBM, timing-only mode; parity on a slice against the oracle.
    assemble_rec = None
    if assemble is not None:
        def extras_for(n, seq_len_, qual_, qoff_):
            import struct
            names = [b"m64011_190830_220126/%d/ccs" % (4194304 + 7 * i) for i in range(n)]
            name_off = np.zeros(n + 1, np.uint64)
            name_off[1:] = np.cumsum([len(x) for x in names])
            aux1 = (b"NMi" + struct.pack("<i", 17) + b"rqf" + struct.pack("<f", 0.999) + b"npi" + struct.pack("<i", 11) + b"ecf" + struct.pack("<f", 10.5)
                    + b"snBf" + struct.pack("<I4f", 4, 9.1, 17.2, 5.3, 9.9) + b"zmi" + struct.pack("<i", 4194304) + b"RGZ" + b"a1b2c3d4" + b"\0")
            aux = np.tile(np.frombuffer(aux1, np.uint8), n)
            aux_off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(len(aux1)))
            return dict(name_off=name_off, names=np.frombuffer(b"".join(names) + b"\0" * 16, np.uint8).copy(), aux_off=aux_off,
                        aux=np.concatenate([aux, np.zeros(16, np.uint8)]), mate_tid=np.full(n, -1, np.int32), mate_pos=np.full(n, -1, np.int32),
                        tlen=np.zeros(n, np.int32), qual=qual_, qual_off=qoff_)
        ctx.set_names(s.contig_names, s.chrom_names)
        ctx.submit_c(ch.c, 0)
        ctx.wait_c(0)
        r0, _ = ctx.assemble_records(extras_for(int(ch.c.n_reads), seq_len, qual, qoff), 0, flags=abi.ASM_NO_DOWNLOAD)
        ms = []
        for _ in range(max(args.warmup, 3) + args.steps):
            o, _ = ctx.assemble_records(None, 0, flags=abi.ASM_RESIDENT_QUAL | abi.ASM_NO_DOWNLOAD)
            ms.append(float(o.kernel_ms))
        r_ms = float(np.mean(ms[max(args.warmup, 3):]))
        r_bytes = int(o.bytes_read + o.bytes_written)
        # the same records framed as level-0 BGZF on the device (the reference's stdout mode), CRC32 in the kernel
        zs = []
        for _ in range(max(args.warmup, 3) + args.steps):
            zo, _ = ctx.bgzf_store_records(b"", 0, flags=abi.ASM_NO_DOWNLOAD)
            zs.append(float(zo.kernel_ms))
        z_ms = float(np.mean(zs[max(args.warmup, 3):]))
        z_bytes = int(zo.bytes_read + zo.bytes_written)
        xs = extras_for(int(sub.c.n_reads), sl_len, sq, so)
        octx_a.set_names(s.contig_names, s.chrom_names)
        t0 = time.perf_counter()
        _, (rbo, byo) = octx_a.assemble_records(xs)
        t_cpu_r = time.perf_counter() - t0
        helpers.lift_c(ctx, sub.c, slot=1)
        _, (rbg, byg) = ctx.assemble_records(xs, 1)
        if not (np.array_equal(rbo, rbg) and np.array_equal(byo, byg)):
            raise SystemExit("PARITY FAILURE vs oracle in ptl_assemble_records on the bench workload")
        import gzip
        _, zb = ctx.bgzf_store_records(b"", 1, flags=abi.BGZF_EOF)
        if gzip.decompress(zb) != byg.tobytes():
            raise SystemExit("ptl_bgzf_store_records: the framed stream does not decompress to the records")
        r_traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["bam_write_kernel"]
            if tj["workload"] == args.workload and tj["reads"] == int(ch.c.n_reads):
                r_traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        assemble_rec = {"kernel": "bam_write_kernel", "what": "every output record as bam_write1 bytes: clone_record tag stripping, field updates, PS/ZM/SA "
                        "tags, bases/qualities re-oriented (src/read_alignment_scanner.rs:105-133,245-282,310-366); inputs resident in HBM",
                        "records": int(o.n_records), "reads": int(ch.c.n_reads), "bam_bytes": int(o.bytes_written), "kernel_ms": r_ms,
                        "records_per_s": o.n_records / (r_ms / 1e3),
                        "roofline": {"bound": "hbm", "achieved": r_bytes / (r_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": r_bytes / (r_ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": r_bytes, "traffic": r_traffic},
                        "cpu_baseline": {"value": (len(rbo) - 1) / t_cpu_r, "unit": "records/s", "cores": 1, "kind": "port",
                                         "sample": f"{sub.c.n_reads} reads of this workload, oracle restatement incl. the python-side copy of the result"},
                        "parity": f"byte-exact vs oracle on {sub.c.n_reads} reads of this workload",
                        "bgzf_store": {"kernel": "bgzf_store_kernel", "what": "the records framed as level-0 BGZF blocks, CRC32 computed on the device "
                                       "(the reference's stdout mode, src/read_alignment_scanner.rs:66-71)", "blocks": int(zo.n_blocks), "kernel_ms": z_ms,
                                       "roofline": {"bound": "hbm", "achieved": z_bytes / (z_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                                    "frac": z_bytes / (z_ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": z_bytes, "traffic": None},
                                       "check": f"python gzip reads the framed stream of {sub.c.n_reads} reads back to the record bytes"}}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_sample = args.cpu_sample or min(n_reads, max(20_000, 2500 * threads))
        sub = lib.PackedBatch(L, s.read_records, 0, n_sample, s.contig_names)
        octx = helpers.oracle_context(s, threads=threads, faithful=True)
        time_oracle(octx, sub.c)
        dt, pairs = time_oracle(octx, sub.c, reps=3)
        octx1 = helpers.oracle_context(s, threads=threads, faithful=False)
        dt_lean, _ = time_oracle(octx1, sub.c, reps=3)
        cpu = {"value": pairs / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n_sample} reads ({pairs} pairs) of this workload, best of 3; oracle = C++ restatement incl. the reference's "
                         f"per-pair read decode; without that decode (lazy base access): {pairs / dt_lean:.3e} {UNIT}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {WORKLOAD_DOC.get(args.workload, 'custom')} ({s.n_chrom} x {int(s.chrom_len[0])} bp reference, "
                                   f"{s.contig_records.n_records} contig alignment records -> {n_segments} segments after trim/join, "
                                   f"{n_reads} reads per GPU)",
                       "reads_per_gpu": int(n_reads), "pairs_per_step_per_gpu": int(cnt["n_pairs"]), "lifted_per_step_per_gpu": int(cnt["n_lifted"]),
                       "l2_policy": "inputs larger than L2 (CIGAR pools + op scratch > 126 MB per step)",
                       "sharding": "one shard per rank (same generator parameters and seed: identical per-GPU work), no collective",
                       "rank_ms_per_step": [round(v, 4) for v in rank_ms],
                       "e2e_pipeline": f"{len(chunks)} batches of {args.chunk} reads over {max(args.slots, 2)} slots", "e2e_base_transfer": best_mode,
                       "parity": parity, "full_batch_digest": f"{list(digests.values())[0]:016x} (identical across {len(digests)} base-transfer modes)",
                       "generate_s": round(t_gen, 1)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes(best_mode)), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_dt * 1e3,
                    "alternatives": {k: {"value": pairs_total / v[0], "ms_per_step": v[0] * 1e3, "h2d_bytes_per_step": int(h2d_bytes(k))} for k, v in e2e_runs.items()}},
            "gpu_launches": int(launches_value),
            "gpu_launches_e2e": int(launches_e2e),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(windows),
            "counters": cnt,
            "assemble_bases": assemble,
            "assemble_records": assemble_rec,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
