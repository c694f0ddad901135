import os, sys, shutil, subprocess, tempfile, random, gzip
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from portello_b200 import bamio, synth, abi
d = tempfile.mkdtemp(prefix="fuzz_")
s = synth.make("tiny", seed=5, n_reads=300)
paths = bamio.write_dataset(s, d, n_unmapped=5)
bam = paths["reads"]
bamio.index_bam_csi(bam, bam[:-4] + ".x.csi", 14, 5)
orig = {k: open(k, "rb").read() for k in (bam, bam + ".bai", bam[:-4] + ".x.csi")}
child = r'''
import sys
sys.path.insert(0, ".")
from portello_b200 import bamio, abi
try:
    f = bamio.BamFile(sys.argv[1])
    if f.has_index:
        for t in range(len(f.ref_names)):
            f.fetch(t, 0, f.ref_len[t]).n
            f.fetch(t, 1000, 50000, bamio.START_IN_REGION).n
        f.fetch(bamio.FETCH_UNMAPPED, flt=bamio.ONLY_UNMAPPED).n
    f.fetch(bamio.FETCH_ALL).n
    print("ok")
except abi.PtlError as e:
    print("err", str(e)[:80])
'''
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bad = 0
work = os.path.join(d, "w.bam")
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 150):
    kind = it % 3
    for f in (work, work + ".bai", work + ".csi"):
        if os.path.exists(f): os.remove(f)
    data = {0: bytearray(orig[bam]), 1: bytearray(orig[bam + ".bai"]), 2: bytearray(orig[bam[:-4] + ".x.csi"])}[kind]
    if kind == 2:   # mutate the decompressed CSI, recompress as BGZF
        raw = bytearray(gzip.decompress(bytes(data)))
        tgt = raw
    else:
        tgt = data
    mode = rng.random()
    if mode < 0.6:
        for _ in range(rng.randint(1, 4)):
            p = rng.randrange(len(tgt)); tgt[p] = rng.randrange(256)
    elif mode < 0.8:
        del tgt[rng.randrange(len(tgt)):]
    else:
        p = rng.randrange(max(1, len(tgt) - 8)); tgt[p:p + 4] = (0xffffffff if rng.random() < 0.5 else 0x7fffffff).to_bytes(4, "little")
    if kind == 2:
        data = bamio.bgzf_compress(bytes(tgt)) if hasattr(bamio, "bgzf_compress") else gzip.compress(bytes(tgt))
    open(work, "wb").write(bytes(data) if kind == 0 else orig[bam])
    if kind == 1: open(work + ".bai", "wb").write(bytes(data))
    elif kind == 2: open(work + ".csi", "wb").write(bytes(data))
    else: open(work + ".bai", "wb").write(orig[bam + ".bai"])
    try:
        r = subprocess.run([sys.executable, "-c", child, work], capture_output=True, text=True, timeout=60)
        out = r.stdout.strip()[:100]
        if r.returncode != 0:
            bad += 1
            print("CRASH", it, kind, r.returncode, r.stderr[-300:].replace("\n", " | "))
    except subprocess.TimeoutExpired:
        bad += 1
        print("HANG", it, kind)
print("done, bad =", bad)
shutil.rmtree(d)
