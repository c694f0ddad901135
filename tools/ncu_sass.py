#!/usr/bin/env python
"""SASS-level view of one kernel of an ncu capture: headline metrics, samples / warp instructions bucketed by the
call-site line in a chosen file, and optionally the per-instruction listing of an address range.
usage: tools/ncu_sass.py <rep> <kernel-substring> [--bucket pair_bodies.cuh] [--range 0x6600 0x6f80] [--top 25]"""
import argparse, collections, csv, io, os, re, subprocess, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("kernel")
ap.add_argument("--bucket", default="pair_bodies.cuh")
ap.add_argument("--range", nargs=2, default=None)
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--so", default=None)
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = a.so or os.path.join(root, "portello_b200", "csrc", "libportello_b200.so")

ksel = ["-k", "regex:" + a.kernel.split("ILb")[0] + "$" if "ILb" not in a.kernel and not a.kernel.endswith("kernel") else "regex:" + a.kernel.split("ILb")[0]]  # (the report may hold several kernels)
raw = subprocess.run(["ncu", "-i", a.rep] + ksel + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:72s} {vals[i]:>16s} {units[i]}")

src = subprocess.run(["ncu", "-i", a.rep] + ksel + ["--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
shdr = rows[1]
ix = {h: i for i, h in enumerate(shdr)}
data = [r for r in rows[2:] if len(r) >= len(shdr) and re.match(r"^(0x)?[0-9a-fA-F]+$", r[ix["Address"]])]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "kernels" in f][0]
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", cub], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and a.kernel in l][0]
end = [i for i, l in enumerate(sass) if i > start and l.startswith(".text.")]
seg = sass[start:(end[0] if end else len(sass))]
cur, fresh, amap = None, True, {}
for l in seg:
    if "//## File" in l:
        m = re.search(r'File "([^"]+)", line (\d+)', l)
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        if fresh:
            cur, fresh = [loc], False
        else:
            cur.append(loc)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        amap[int(m.group(1), 16)] = cur[:] if cur else None
        fresh = True


def addr(r):
    x = r[ix["Address"]]
    return int(x, 16) if x.startswith("0x") else int(x)


seen_addr, uniq = set(), []
for r in data:  # (a filtered multi-kernel report lists the kernel's instructions once per matching result)
    if r[ix["Address"]] not in seen_addr:
        seen_addr.add(r[ix["Address"]])
        uniq.append(r)
data = uniq
base = addr(data[0])
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"total samples {tot_s}, warp instructions {tot_i / 1e6:.1f}M, {len(data)} SASS instructions")
b = collections.defaultdict(lambda: [0, 0, 0])
stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
sb = collections.defaultdict(lambda: collections.Counter())
for r in data:
    ad = addr(r) - base
    loc = amap.get(ad)
    key = "?"
    if loc:
        pb = [f for f in loc if f[0] == a.bucket]
        key = f"{a.bucket}:{pb[0][1]}" if pb else f"{loc[-1][0]}:{loc[-1][1]}"
    s, i, t = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
    b[key][0] += s; b[key][1] += i; b[key][2] += t
    for h in stalls:
        v = int(r[ix[h]])
        if v:
            sb[key][h] += v
print(f"== by call-site line in {a.bucket}")
for k, (s, i, t) in sorted(b.items(), key=lambda x: -x[1][0])[: a.top]:
    top = ", ".join(f"{h[6:]} {v}" for h, v in sb[k].most_common(4))
    print(f"{k:28s} samples {100 * s / tot_s:5.1f}%  inst {i / 1e6:7.2f}M ({100 * i / tot_i:4.1f}%)  thr/inst {t / max(i, 1):4.1f}   {top}")
if a.range:
    lo, hi = int(a.range[0], 16), int(a.range[1], 16)
    print("== per instruction")
    for r in data:
        ad = addr(r) - base
        if lo <= ad < hi:
            s, i, t = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
            st = sorted(((h[6:], int(r[ix[h]])) for h in stalls if int(r[ix[h]]) > 0), key=lambda x: -x[1])[:2]
            loc = amap.get(ad)
            print("%04x %5.2f%% i=%5.2fM t/i=%4.1f %-16s %-52s %s" % (ad, 100 * s / tot_s, i / 1e6, t / max(i, 1), (f"{loc[0][0][:10]}:{loc[0][1]}" if loc else ""), r[ix["Source"]][:52], st))
