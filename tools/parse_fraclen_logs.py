#!/usr/bin/env python
"""Parses the per-layer (input_fraclen, weight_fraclen) dumps the reference printed after its
own trained runs (fix_train.py:970-991; logs shipped under /root/reference/fraclen_visual/) into
f8net_b200/data/trained_fraclens_<name>.json -- the "trained-fraclen" fixture family of
SURVEY.md 8(d): real per-layer formats (fi 1..8, fw 0..7, including the fw in {0,1} layers)
instead of the calibrated synthetic ones.  Authoring container only (needs /root/reference).

The logs name FLOAT-model layers (body.0, body.1, body.2: no ReLU modules between the convs);
the int model's prefixes are body.0 / body.2 / body.4 (export.float_layers holds the mapping).
The last dump of each log is used (the state after training), parsed like the reference's own
fraclen_visualizing_{mbv2,res50}.py: input_fraclen is rounded, weight_fraclen taken as is.
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from f8net_b200.arch import graph_for  # noqa: E402
from f8net_b200.export import ExportFlags, float_layers  # noqa: E402

LOGS = {
    "mobilenet_v2": ("mobilenet_v2", False, "mbv2_fix_quant.out"),
    "resnet50_ptcv": ("resnet50", True, "res50_fix_quant_ptcv_pretrained.out"),
    "resnet50_nvidia": ("resnet50", True, "res50_fix_quant_nvidia_pretrained.out"),
}
REF = "/root/reference/fraclen_visual"


def parse(path):
    """All 'layer name / input_fraclen / weight_fraclen' triples; later dumps overwrite earlier ones."""
    table, name = {}, None
    fi = None
    for line in open(path):
        line = line.strip()
        m = re.match(r"layer name: (\S+?)\.$", line)
        if m:
            name, fi = m.group(1), None
            continue
        m = re.match(r"input_fraclen: tensor\(\[([-0-9.e]+)\]", line)
        if m and name:
            fi = int(round(float(m.group(1))))
            continue
        m = re.match(r"weight_fraclen: ([-0-9.e]+)\.$", line)
        if m and name and fi is not None:
            table[name] = (fi, int(round(float(m.group(1)))))
            name = None
    return table


def main():
    out_dir = os.path.join(ROOT, "f8net_b200", "data")
    for key, (arch, normalize, fname) in LOGS.items():
        src = os.path.join(REF, fname)
        tab = parse(src)
        net = graph_for(arch, normalize)
        res, hist_i, hist_w = {}, {}, {}
        for L in float_layers(net, ExportFlags(normalize=normalize)):
            fi, fw = tab[L.fprefix]
            res[L.iprefix] = [fi, fw]
            hist_i[fi] = hist_i.get(fi, 0) + 1
            hist_w[fw] = hist_w.get(fw, 0) + 1
        assert len(res) == len(tab) == len(net.convs()), (len(res), len(tab))
        doc = {"source": f"fraclen_visual/{fname} (last per-layer dump)", "arch": arch,
               "head_signed": normalize, "fraclens": res}
        with open(os.path.join(out_dir, f"trained_fraclens_{key}.json"), "w") as f:
            json.dump(doc, f, indent=0, sort_keys=True)
        print(key, len(res), "layers; fi", dict(sorted(hist_i.items())), "fw", dict(sorted(hist_w.items())))


if __name__ == "__main__":
    main()
