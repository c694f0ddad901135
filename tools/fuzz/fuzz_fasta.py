import sys, random, os, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from portello_b200 import bamio, abi
rng = random.Random(5)
d = tempfile.mkdtemp()
base = b">chrA first one\nacgtNNryACGT\r\nACGT\n>chrB\n\n>chrC\tx\nttt\n>d\n" + b"ACGT" * 5000 + b"\n"
n_ok = n_err = 0
for it in range(600):
    t = bytearray(base)
    for _ in range(rng.randint(1, 6)):
        m = rng.random()
        if m < 0.5: t[rng.randrange(len(t))] = rng.randrange(256)
        elif m < 0.7: del t[rng.randrange(len(t)):]
        elif m < 0.85: t[rng.randrange(len(t)):rng.randrange(len(t))] = b">" 
        else: t = bytearray(b"\n" * rng.randint(0, 3)) + t
    p = os.path.join(d, "x.fa"); open(p, "wb").write(bytes(t))
    try:
        bamio.load_fasta(p, threads=rng.choice([1, 3])); n_ok += 1
    except abi.PtlError:
        n_err += 1
print("done", n_ok, n_err)
