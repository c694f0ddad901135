#!/usr/bin/env python
"""Attribute an ncu capture of one kernel to source lines (SASS page x nvdisasm line info), and print headline metrics.
usage: tools/ncu_lines.py gpurun_out/prof_X.ncu-rep <kernel-substring> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "portello_b200", "csrc", "libportello_b200.so")
KFILT = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []
raw = subprocess.run(["ncu", "-i", rep, *KFILT, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "launch__grid_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed_op_local_ld.sum", "lts__t_bytes.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:72s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, *KFILT, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
second = next((i for i, r in enumerate(rows[1:], 1) if r and r[0] == "Kernel Name"), None)
if second:
    rows = rows[:second]  # (several launches of the kernel in the capture: the first one)
shdr = rows[1]
ix = {h: i for i, h in enumerate(shdr)}
data = rows[2:]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "kernels" in f][0]
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", cub], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
start = end = None
for i, l in enumerate(sass):
    if l.startswith(".text.") and kname in l:
        start = i
    elif start is not None and l.startswith(".text.") and kname not in l:
        end = i
        break
seg = sass[start:end]
cur, insts, fresh = None, [], True
for l in seg:
    if "//## File" in l:
        m = re.search(r'File "([^"]+)", line (\d+)', l)
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        if fresh:
            cur = [loc]      # innermost frame first
            fresh = False
        else:
            cur.append(loc)  # enclosing (inlined-at) frames follow
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((m.group(2), tuple(cur) if cur else None))
        fresh = True
assert len(insts) == len(data), (len(insts), len(data))
def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0
for level, name in ((0, "innermost line"), (-1, "outermost (kernel) line")):
    agg, ai, at = collections.Counter(), collections.Counter(), collections.Counter()
    for r, (txt, loc) in zip(data, insts):
        k = loc[level] if loc else None
        agg[k] += f(r, "# Samples"); ai[k] += f(r, "Instructions Executed"); at[k] += f(r, "Thread Instructions Executed")
    tot = sum(agg.values()) or 1
    print(f"\n== by {name}: total samples {tot:.0f}, warp insts {sum(ai.values())/1e6:.1f}M")
    for k, v in agg.most_common(top):
        print(f"{v/tot*100:5.1f}%  inst {ai[k]/1e6:7.2f}M  thr/inst {at[k]/max(ai[k],1):5.1f}  {k}")
stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        tot[s] += f(r, s)
print("\n== stall reasons:", [(k, int(v)) for k, v in tot.most_common(6)])
