#!/usr/bin/env python
"""File-to-file throughput of the command-line tool (under gpurun): write a synthetic data set as indexed BAMs + FASTA, run
portello-b200 on it in file mode and in pipe mode, report reads/s and GB/s of BAM in/out next to the stage the GPU work is.
usage: python tools/cli_e2e.py [workload] [n_reads] [threads]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from portello_b200 import bamio, lib, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "chr20"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 4)
d = tempfile.mkdtemp(prefix="ptl_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
t0 = time.time()
s = synth.make(wl, n_reads=n)
paths = bamio.write_dataset(s, d, n_unmapped=n // 200, level=1, threads=threads)
t_write = time.time() - t0
in_bytes = os.path.getsize(paths["reads"])
cli = os.path.join(lib.CSRC, "portello-b200")
out = {"workload": wl, "reads": n, "threads": threads, "input_bam_bytes": in_bytes, "dataset_write_s": round(t_write, 1), "runs": {}}
for mode in ("file", "stdout"):
    o, u = os.path.join(d, f"remapped_{mode}.bam"), os.path.join(d, f"unassembled_{mode}.bam")
    args = [cli, "--assembly-to-ref", paths["contigs"], "--read-to-assembly", paths["reads"], "--ref", paths["ref"],
            "--remapped-read-output", o if mode == "file" else "-", "--unassembled-read-output", u, "--threads", str(threads)]
    t0 = time.time()
    if mode == "file":
        r = subprocess.run(args, capture_output=True)
    else:
        with open(o, "wb") as fh:
            r = subprocess.run(args, stdout=fh, stderr=subprocess.PIPE)
    dt = time.time() - t0
    if r.returncode != 0:
        print(r.stderr.decode()[-2000:])
        raise SystemExit(1)
    out["runs"][mode] = {"wall_s": round(dt, 2), "reads_per_s": round(n / dt), "output_bam_bytes": os.path.getsize(o),
                         "summary": [l for l in r.stderr.decode().splitlines() if "Lifted" in l][-1].split("] ", 1)[-1]}
    if os.environ.get("CLI_E2E_LOG"):
        print(f"--- {mode} log", file=sys.stderr)
        print(r.stderr.decode(), file=sys.stderr)
subprocess.run(["rm", "-rf", d])
print(json.dumps(out))
