import os, sys, random
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests'))
import numpy as np
import test_host_logic as T
from portello_b200 import lib, abi
L = lib.load()
rng = random.Random(3)
valid = ["ctg1,100,-,20S30=1I9=,60,0;", "ctg0,900,+,1I39=20S,33,1;ctg1,5,-,40S3=1I16=,60,0;", "ctg1,1,+,60M,0,0;"]
alphabet = "0123456789,;+-MIDNSHP=X ctg\t*"
n_ok = n_err = 0
for it in range(4000):
    sa = list(rng.choice(valid))
    for _ in range(rng.randint(0, 4)):
        m = rng.random()
        if m < 0.4 and sa: sa[rng.randrange(len(sa))] = rng.choice(alphabet)
        elif m < 0.6 and sa: del sa[rng.randrange(len(sa))]
        elif m < 0.8: sa.insert(rng.randrange(len(sa) + 1), rng.choice(alphabet))
        else: sa = sa + list(rng.choice(valid))
    sa = "".join(sa).replace("\0", "")
    row = (rng.choice([0, 16]), 500, 0, "20=40S", 60, sa)
    try:
        rc = T._records([row])
        pb = lib.PackedBatch(L, rc, 0, 1, ["ctg0", "ctg1"], windows=True)
        n_ok += 1
    except abi.PtlError:
        n_err += 1
print("done", n_ok, n_err)
