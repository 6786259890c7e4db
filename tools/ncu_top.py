#!/usr/bin/env python
"""Top stalled SASS instructions per kernel from `ncu --page source --csv` output.
usage: ncu_top.py file.csv [section index] [top n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
print("sections:", len(heads))
h = heads[sec]
end = heads[sec + 1] - 1 if sec + 1 < len(heads) else len(rows)
print(rows[h - 1][:2])
hdr = rows[h]
data = [r for r in rows[h + 1:end] if len(r) == len(hdr)]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[i_samp] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[i_samp] or 0))[:n]:
    st = {hdr[i]: int(r[i] or 0) for i in stall_cols if int(r[i] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(r[i_samp].rjust(6), r[i_exec].rjust(8), r[i_src].strip()[:72].ljust(72), st)
