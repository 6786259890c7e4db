#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv` export: python tools/src_hot.py file.csv [top]
Prints, per SASS instruction with the most stall samples, the samples and the dominant stall reasons."""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(v):
        try:
            return float(v)
        except ValueError:
            return 0.0

    tot = sum(num(r[ix["# Samples"]]) for r in data)
    tot_inst = sum(num(r[ix["Instructions Executed"]]) for r in data)
    print(f"{len(data)} instructions, {tot:.0f} samples, {tot_inst:.0f} warp instructions executed")
    agg = {c: sum(num(r[ix[c]]) for r in data) for c in stall_cols}
    print("stall totals:", ", ".join(f"{c[6:]} {v / tot * 100:.1f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v / tot > 0.005))
    order = sorted(range(len(data)), key=lambda i: -num(data[i][ix["# Samples"]]))[:top]
    for i in sorted(order):
        r = data[i]
        s = num(r[ix["# Samples"]])
        reasons = sorted(((num(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:3]
        rs = " ".join(f"{n}:{v:.0f}" for v, n in reasons if v > 0)
        print(f"{i:5d} {s / tot * 100:5.2f}% exec {num(r[ix['Instructions Executed']]):9.0f}  {r[ix['Source']][:70]:70s} {rs}")


if __name__ == "__main__":
    main()
