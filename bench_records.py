"""Record assembly ("next" rows of SURVEY.md §8f) measured next to the headline metric: bench.py calls measure() on rank 0
at N = 1, outside every timed region of the liftover metric.  One chunk of the workload, bases + qualities + names + aux
resident in HBM, the assembly kernels in their timing-only mode (no H2D, no D2H); parity of the same call path on a slice
against the oracle (checker only)."""
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def measure(args, ctx, s, L, ch, peak):
    import helpers
    from portello_b200 import abi, lib

    n_reads = s.read_records.n_reads
    rank, world = 0, 1
    # ---------------------------------------------------------------- next row: record assembly, bases
    # Outside every timed region of the headline metric.  One chunk of the workload, bases + qualities resident in HBM,
    # ptl_assemble_bases in its timing-only mode (no H2D, no D2H): seq revcomp + qual reverse of every output record.
    assemble = None
    if rank == 0 and world == 1 and not args.no_assemble:
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
        # the same in ONE pass (ptl_frame_records): records never materialised, every byte moved once
        fs = []
        for _ in range(max(args.warmup, 3) + args.steps):
            fo, _ = ctx.frame_records(None, b"", 0, flags=abi.ASM_RESIDENT_QUAL | abi.ASM_NO_DOWNLOAD)
            fs.append(float(fo.kernel_ms))
        f_ms = float(np.mean(fs[max(args.warmup, 3):]))
        f_bytes = int(fo.bytes_read + fo.bytes_written)
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
        helpers.lift_c(ctx, sub.c, slot=1)
        _, zf = ctx.frame_records(xs, b"", 1, flags=abi.BGZF_EOF)
        if zf != zb:
            raise SystemExit("ptl_frame_records differs from ptl_assemble_records + ptl_bgzf_store_records on the bench workload")
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
                                       "check": f"python gzip reads the framed stream of {sub.c.n_reads} reads back to the record bytes"},
                        "frame_records": {"kernel": "bam_write_meta_kernel + bam_frame_kernel", "what": "ptl_frame_records: assembly fused with the level-0 framing; "
                                          "the payload of every BGZF block straight from the sources, CRC32 from the block's own output (L2)",
                                          "kernel_ms": f_ms, "vs_two_pass_ms": r_ms + z_ms, "records_per_s": o.n_records / (f_ms / 1e3),
                                          "roofline": {"bound": "hbm", "achieved": f_bytes / (f_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                                       "frac": f_bytes / (f_ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": f_bytes, "traffic": None},
                                          "check": "byte-identical to assemble + store on this slice"}}


    return {"assemble_bases": assemble, "assemble_records": assemble_rec, "e2e_records": measure_e2e_records(args, ctx, s, L, extras_for)}


def pinned_array(L, nbytes):
    import ctypes as C
    p = L.dll.ptl_host_alloc(max(int(nbytes), 64))
    if not p:
        raise MemoryError("ptl_host_alloc failed")
    return np.ctypeslib.as_array((C.c_uint8 * int(nbytes)).from_address(p)), p


def measure_e2e_records(args, ctx, s, L, extras_for, n_chunks=6, chunk=65536):
    """The record path end to end (src/read_alignment_scanner.rs:61-78,105-133,245-282,482-487): DECODED records (packed bases,
    qualities, names, aux) in pinned host memory -> H2D -> lift -> record assembly -> level-0 BGZF framing + CRC on the
    device -> D2H of the framed bytes, every copy inside the timed region; three host threads, one per batch slot."""
    import threading
    from portello_b200 import abi, lib

    n_reads = s.read_records.n_reads
    chunk = min(chunk, n_reads)
    n_chunks = max(1, min(n_chunks, n_reads // chunk))
    ctx.set_seq_zero_copy(False)  # the assembly streams the bases from HBM: they are uploaded with the batch
    packs, extras, keep = [], [], []
    tile = np.random.default_rng(1).integers(0, 94, 1 << 24, dtype=np.uint8)
    in_bytes = 0
    for k in range(n_chunks):
        pb = lib.PackedBatch(L, s.read_records, k * chunk, chunk, s.contig_names, pinned=True)
        seq_len = np.ctypeslib.as_array(pb.c.read_seq_len, (pb.c.n_reads,)).astype(np.int64)
        qoff = np.zeros(pb.c.n_reads, np.uint64)
        qoff[1:] = np.cumsum(seq_len[:-1])
        x = extras_for(int(pb.c.n_reads), seq_len, None, qoff)
        qual, qp = pinned_array(L, int(seq_len.sum()) + 64)
        qual[:] = np.resize(tile, qual.size)
        x["qual"] = qual
        for f in ("names", "aux"):  # pinned copies of the pools
            a, ap = pinned_array(L, x[f].size)
            a[:] = x[f]
            x[f] = a
            keep.append(ap)
        keep.append(qp)
        packs.append(pb)
        extras.append(x)
        in_bytes += int(pb.c.seq4_bytes) + qual.size + x["names"].size + x["aux"].size + int(pb.c.n_cigar) * 4 + int(pb.c.n_reads) * 40
    ctx.set_names(s.contig_names, s.chrom_names)
    out_bytes = [0] * n_chunks
    n_records = [0] * n_chunks
    errors = []

    def worker(slot, reps):
        try:
            for _ in range(reps):
                for k in range(slot, n_chunks, 3):
                    ctx.submit_c(packs[k].c, slot)
                    ctx.wait_c(slot)
                    o, _ = ctx.assemble_records(extras[k], slot, flags=abi.ASM_NO_DOWNLOAD)
                    z, zb = ctx.bgzf_store_records(b"", slot, flags=0, copy=False)
                    out_bytes[k] = int(z.n_bytes)
                    n_records[k] = int(o.n_records)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def run(reps):
        th = [threading.Thread(target=worker, args=(sl, reps)) for sl in range(3)]
        [t.start() for t in th]
        [t.join() for t in th]
        if errors:
            raise errors[0]

    run(1)
    reps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    run(reps)
    dt = (time.perf_counter() - t0) / reps
    for p in keep:
        L.dll.ptl_host_free(p)
    recs = sum(n_records)
    return {"what": "decoded records (packed bases, qualities, names, aux) in pinned host memory -> lift -> ptl_assemble_records -> "
                    "ptl_bgzf_store_records -> level-0 BGZF bytes on the host; H2D and D2H inside the timed region, one host thread per slot",
            "reads_per_pass": n_chunks * chunk, "records_per_pass": recs, "ms_per_pass": dt * 1e3, "records_per_s": recs / dt,
            "h2d_bytes_per_pass": in_bytes, "d2h_bytes_per_pass": sum(out_bytes),
            "pcie_gbs": {"h2d": round(in_bytes / dt / 1e9, 2), "d2h": round(sum(out_bytes) / dt / 1e9, 2)},
            "bound": "PCIe: ~23 KB in and ~23 KB out per read; the kernels of a pass take a few ms"}
