#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libportello_b200.so the instruction count and the histogram of its memory
instructions by width (what the streaming kernels actually issue), plus the full listing of the kernels named on the
command line.  usage: tools/sass_summary.py [kernel-substring ...] > profiles/rNN_sass.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "portello_b200", "csrc", "libportello_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
kernels, cur = collections.OrderedDict(), None
for l in sass:
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "")
        kernels[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        kernels[cur].append(l.rstrip())
MEM = re.compile(r"\b(LDG|STG|LD|ST|LDS|STS|LDSM|LDL|STL|ATOMG|ATOMS|RED|LDC|LDGSTS|UBLKCP|SHFL|PRMT|BAR)(\.[A-Z0-9_.]+)?\b")
print(f"# {os.path.relpath(so, ROOT)}: {len(kernels)} kernels (cuobjdump -sass, sm_100a)\n")
for k, ins in kernels.items():
    h = collections.Counter()
    for l in ins:
        body = l.split("*/", 1)[1]
        m = MEM.search(body)
        if m:
            op = m.group(1) + (m.group(2) or "")
            op = re.sub(r"\.(E|CONSTANT|STRONG|SYS|GPU|SM|CTA|LTC128B|EF|EL|U|SYNC|DEFER_BLOCKING|BFLY|UP|DOWN|IDX|MMIO)\b", "", op)
            h[op] += 1
    mem = ", ".join(f"{o} x{c}" for o, c in sorted(h.items()) if not o.startswith(("PRMT", "SHFL", "BAR", "LDC")))
    oth = ", ".join(f"{o} x{c}" for o, c in sorted(h.items()) if o.startswith(("PRMT", "SHFL", "BAR")))
    print(f"{k}: {len(ins)} instructions\n    memory: {mem}\n    other:  {oth}\n")
for want in sys.argv[1:]:
    for k, ins in kernels.items():
        if want in k:
            print(f"\n## full listing: {k}\n")
            print("\n".join(ins))
