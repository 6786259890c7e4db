#!/usr/bin/env python
"""Turns the ncu artefacts of tools/collect_profiles.sh into the committed summaries under profiles/.

    python tools/make_profiles.py r02 [--src gpurun_out] [--dst profiles]

Per network (resnet18, resnet50, mobilenet_v1, mobilenet_v2), from ONE forward pass at batch 256:
  <tag>_launches_<arch>.md   per-launch gpu__time_duration (ncu --metrics, cold caches, serialised: use the SHARES)
                             grouped by kernel template, beside the algorithmic bytes / int8 ops of each launch
  <tag>_ncu_full_<arch>.md   one row per launch from `ncu --set full`: DRAM bytes, DRAM / L2 / tensor-pipe / issue
                             utilisation, registers, shared memory
  r02_traffic.json           DRAM bytes per launch per kernel template (bench.py's roofline.traffic)
The launches of the capture are matched to plan ops by order (tools/one_pass.py --names).
"""
import argparse
import csv
import json
import os
import re

ARCHS = ["resnet18", "resnet50", "mobilenet_v1", "mobilenet_v2"]


def short(name):
    name = re.sub(r"void |<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    return name.split("(")[0].strip()


def read_csv(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    return rows[0], rows[1:]


def num(v):
    try:
        return float(v.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def launches(src, dst, tag, arch):
    path = os.path.join(src, f"{tag}_launches_{arch}.csv")
    names = json.load(open(os.path.join(src, f"{tag}_names_{arch}.json")))["launches"]
    hdr, rows = read_csv(path)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    rows = [r for r in rows if len(r) == len(hdr)]
    if len(rows) != len(names):
        print(f"  {arch}: {len(rows)} captured launches vs {len(names)} plan launches -- matched by order up to the shorter")
    tmpl = {}
    total = 0.0
    per = []
    for r, n in zip(rows, names):
        key = n["kernel"].split("<")[0]
        if key not in r[ik]:
            print(f"  {arch}: launch order mismatch: plan says {n['kernel']}, ncu captured {short(r[ik])}")
        us = num(r[iv]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
        total += us
        t = tmpl.setdefault(n["kernel"], [0, 0.0, 0.0, 0.0])
        t[0] += 1
        t[1] += us
        t[2] += n["algorithmic_bytes"]
        t[3] += n["int8_ops"]
        per.append((n["op"], n["kernel"], short(r[ik]), us, n["algorithmic_bytes"], n["int8_ops"]))
    out = os.path.join(dst, f"{tag}_launches_{arch}.md")
    with open(out, "w") as f:
        f.write(f"# {arch}, batch 256, one forward pass under `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n"
                f"Cold-cache, serialised launches: the SHARES are meaningful, the absolutes are not (sum {total:.0f} us).\n\n"
                "| kernel template | launches | total us | share | algorithmic GB/s | int8 TOPS |\n|---|---|---|---|---|---|\n")
        for k, (n, us, b, ops) in sorted(tmpl.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f} % | {b / us / 1e3:.0f} | {ops / us / 1e6:.0f} |\n")
        f.write("\n| # | op | kernel template | ncu kernel name | us |\n|---|---|---|---|---|\n")
        for i, (op, k, nk, us, _, _) in enumerate(per):
            f.write(f"| {i} | {op} | `{k}` | `{nk[:60]}` | {us:.1f} |\n")
    print("wrote", out)


def full(src, dst, tag, arch, traffic):
    path = os.path.join(src, f"{tag}_full_{arch}.csv")
    names = json.load(open(os.path.join(src, f"{tag}_names_{arch}.json")))["launches"]
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, name, fmt="{:.1f}", scale=1.0):
        v = num(r[ix[name]]) if name in ix else None
        return "-" if v is None else fmt.format(v * scale)

    def bytes_of(r, name):
        if name not in ix:
            return 0.0
        v = num(r[ix[name]])
        mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(units[ix[name]], 1.0)
        return (v or 0.0) * mul

    out = os.path.join(dst, f"{tag}_ncu_full_{arch}.md")
    fam = {}
    with open(out, "w") as f:
        f.write(f"# {arch}, batch 256, one forward pass under `ncu --set full --clock-control none` (final round-2 binary)\n\n"
                "| # | op | kernel template | grid | us | DRAM rd MB | DRAM wr MB | alg MB | DRAM % | L2 % | tensor % | issue % | warps % | regs | smem KB |\n"
                "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for i, (r, n) in enumerate(zip(data, names)):
            rd, wr = bytes_of(r, "dram__bytes_read.sum"), bytes_of(r, "dram__bytes_write.sum")
            e = fam.setdefault(n["kernel"], [0, 0.0])
            e[0] += 1
            e[1] += rd + wr
            dur = num(r[ix["gpu__time_duration.sum"]]) or 0.0
            dur *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1e-3)
            f.write(f"| {i} | {n['op']} | `{n['kernel']}` | {col(r, 'launch__grid_size', '{:.0f}')} | {dur:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                    f"{n['algorithmic_bytes'] / 1e6:.1f} | {col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | "
                    f"{col(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                    f"{col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | "
                    f"{col(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active')} | "
                    f"{col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                    f"{col(r, 'launch__registers_per_thread', '{:.0f}')} | "
                    f"{col(r, 'launch__shared_mem_per_block_dynamic', '{:.0f}', 1e-3 if units[ix.get('launch__shared_mem_per_block_dynamic', 0)] == 'byte/block' else 1.0)} |\n")
    traffic[arch] = {"batch": 256, "source": f"profiles/{tag}_ncu_full_{arch}.md (ncu --set full, one forward pass, final binary)"}
    for k, (n, b) in fam.items():
        traffic[arch][k] = {"launches": n, "dram_bytes_per_launch": b / n}
    print("wrote", out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--src", default="gpurun_out")
    ap.add_argument("--dst", default="profiles")
    a = ap.parse_args()
    tpath = os.path.join(a.dst, "r02_traffic.json")
    try:
        traffic = json.load(open(tpath))
    except (OSError, ValueError):
        traffic = {}
    for arch in ARCHS:
        if os.path.exists(os.path.join(a.src, f"{a.tag}_launches_{arch}.csv")):
            launches(a.src, a.dst, a.tag, arch)
        if os.path.exists(os.path.join(a.src, f"{a.tag}_full_{arch}.csv")):
            full(a.src, a.dst, a.tag, arch, traffic)
    json.dump(traffic, open(tpath, "w"), indent=1)
    print("updated", tpath)


if __name__ == "__main__":
    main()
