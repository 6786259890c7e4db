#!/usr/bin/env python
"""SASS opcode histogram of the shipped library: which Blackwell instructions it contains (tcgen05.mma =
UTCIMMA, tcgen05.ld = LDTM, TMA = UTMALDG / UTMASTG, bulk copies = UBLKCP, mbarriers = SYNCS, legacy IMMA,
dp4a = IDP.4A), for the whole library and per kernel.   python tools/sass_histogram.py [lib] [out]"""
import collections
import datetime
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "f8net_b200/libf8b200.so"
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_sass_histogram.txt"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCIMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG\.\dD|UTMASTG\.\dD|UTMAPF|UBLKCP|SYNCS\.[A-Z0-9.]+|IMMA\.\d+|ELECT|"
                 r"ACQBULK|BAR\.(?:SYNC|ARV)|FENCE\.VIEW\.ASYNC\.[A-Z]+|IDP\.4A\.[A-Z0-9.]+|REDUX|SHFL\.[A-Z]+)")
whole = collections.Counter()
keys = [("utcimma", "UTCIMMA"), ("ldtm", "LDTM"), ("tma_ld", "UTMALDG"), ("tma_st", "UTMASTG"), ("bulk", "UBLKCP"),
        ("imma", "IMMA."), ("dp4a", "IDP.4A")]
per, name = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    m = pat.search(line)
    if m:
        whole[m.group(1)] += 1
    if name:
        for k, p in keys:
            if p in line:
                per[name][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
with open(out, "w") as f:
    f.write(f"# SASS opcode histogram of {lib} ({datetime.date.today()}), from cuobjdump -sass\n# whole library\n")
    for k, v in whole.most_common():
        f.write(f"{v:7d} {k}\n")
    f.write("\n# per kernel\n")
    for n, c in zip(names, per.values()):
        n = re.sub(r"\(anonymous namespace\)::|void ", "", n)
        n = re.sub(r"\(.*$", "", n)
        f.write(f"{n:78s} " + " ".join(f"{k}={c[k]}" for k, _ in keys) + "\n")
print(out, len(per), "kernels")
