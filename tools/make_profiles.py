#!/usr/bin/env python
"""Turns the ncu artefacts of one gpurun call into the committed summaries under profiles/.

    python tools/make_profiles.py launches <launches.csv> <out.md> <title>
    python tools/make_profiles.py full <rep.ncu-rep> <out.md> <title> [traffic.json arch batch]

`launches`: per-kernel totals and shares from `ncu --metrics gpu__time_duration.sum --csv`.
`full`: one row per captured launch from `ncu --set full` (read with `ncu -i ... --page raw --csv`),
and optionally the DRAM bytes per launch per kernel family for bench.py's roofline.traffic.
"""
import csv
import json
import os
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"void |<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    return name.split("(")[0].strip()


def ours(name):
    return any(k in name for k in ("umma_kernel", "head_pool", "dw3x3", "pool_requant", "convert_input",
                                   "maxpool_kernel", "requant_i32", "conv_mma", "pool_fc"))


def family(name):
    if "head_pool" in name:
        return "head_conv_pool"
    if "dw3x3" in name or ", 1>" in name.replace(" ", "")[-6:] and "conv3x3" in name and False:
        return "conv_dw3x3"
    if "umma_kernel" in name or "conv_mma" in name:
        return "conv_dense"
    if "pool_fc" in name:
        return "pool_fc"
    if "pool_requant" in name:
        return "pool_requant"
    if "convert_input" in name:
        return "convert_input"
    return "other"


def launches(path, out, title):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = {}
    for r in rows[1:]:
        if len(r) != len(hdr) or not ours(r[ik]):
            continue
        k = short(r[ik])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r[iv].replace(",", "")) / 1e3      # ns -> us
    total = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\n| kernel | launches | total us | share of our kernels |\n|---|---|---|---|\n")
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f} % |\n")
    print(f"wrote {out}: {len(tot)} kernels, {total:.0f} us")


def full(rep, out, title, traffic=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[0], rows[2:]
    if os.environ.get("F8_ROWS"):          # keep exactly one forward pass when the capture window overlaps two
        data = data[:int(os.environ["F8_ROWS"])]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, name, scale=1.0, fmt="{:.1f}"):
        try:
            return fmt.format(float(r[ix[name]].replace(",", "")) * scale)
        except (KeyError, ValueError):
            return "-"

    fam = {}
    with open(out, "w") as f:
        f.write(f"# {title}\n\n")
        f.write("| # | kernel | grid | time us | dram rd MB | dram wr MB | dram % | L2 % | tensor % | warps % | regs | smem KB |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for n, r in enumerate(data):
            name = short(r[ix["Kernel Name"]])
            f.write(f"| {n} | {name} | {col(r, 'launch__grid_size', fmt='{:.0f}')} | {col(r, 'gpu__time_duration.sum')} | "
                    f"{col(r, 'dram__bytes_read.sum')} | {col(r, 'dram__bytes_write.sum')} | "
                    f"{col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | "
                    f"{col(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                    f"{col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | "
                    f"{col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                    f"{col(r, 'launch__registers_per_thread', fmt='{:.0f}')} | "
                    f"{col(r, 'launch__shared_mem_per_block_dynamic')} |\n")
            try:
                units_rd = rows[1][ix["dram__bytes_read.sum"]]
                mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(units_rd, 1e6)
                b = (float(r[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]])) * mul
                e = fam.setdefault(family(r[ix["Kernel Name"]]), [0, 0.0])
                e[0] += 1
                e[1] += b
            except (KeyError, ValueError):
                pass
    print(f"wrote {out}: {len(data)} launches")
    if traffic:
        path, arch, batch = traffic
        try:
            cur = json.load(open(path))
        except (OSError, ValueError):
            cur = {}
        cur[arch] = {"batch": int(batch), "source": f"{out} (ncu --set full, one forward pass)"}
        for k, (n, b) in fam.items():
            cur[arch][k] = {"launches": n, "dram_bytes_per_launch": b / n}
        json.dump(cur, open(path, "w"), indent=1)
        print(f"updated {path}[{arch}]")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5:8] if len(sys.argv) >= 8 else None)
