#!/usr/bin/env python
"""the reference's own GPU inference mode -- ``integize`` -- beside the engine.

In ``integize`` mode (FLAGS.integize, fix_train.py:895-947; IntModel / IntBlock non-int_op_only
branches: fix_resnet.py:78-118, 384-409, fix_mobilenet_v1.py:39-49, 148-166,
fix_mobilenet_v2.py:49-78, 242-270) the reference keeps the exported INTEGERS in float32 tensors
and runs them through cuDNN / cuBLAS: per int layer ``fix_quant(x, 8, fi, 1, sym) * 2^fi`` ->
float conv with float(int_weight), float(int_bias) -> ``div_(2^(fw+fi))``; residual adds on the
real-valued tensors rescaled to the common fraclen.  The reference is not present on the GPU box,
so this file RESTATES that forward in plain PyTorch from the same IntModel state_dict (it is a
measurement baseline, not product code: nothing under f8net_b200/ imports it) and

  * times it on the B200 (CUDA events, eager and CUDA-graph replay, TF32 off as the float path
    needs exact fp32 products) next to the engine on the same weights and inputs;
  * counts where its logits differ from the exact integer path -- float32 accumulation stops being
    exact once a partial sum passes 2^24, which is why int_op_only exists.

    python tools/integize_gpu.py --arch resnet18 --batch 256
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import f8net_b200  # noqa: E402
from f8net_b200 import synth  # noqa: E402
from f8net_b200.arch import graph_for  # noqa: E402


class Layer:
    def __init__(self, sd, spec, dev):
        p = spec.prefix
        self.spec = spec
        self.w = torch.from_numpy(sd[p + ".weight"]).to(dev).float()
        self.b = torch.from_numpy(sd[p + ".bias"]).to(dev).float()
        self.fw = int(sd[p + ".weight_fraclen"].reshape(-1)[0])
        self.fi = int(sd[p + ".input_fraclen"].reshape(-1)[0])

    def quant_in(self, x):
        """(fix_quant(x, 8, fi, 1, sym)[0] * 2^fi).int().float()  (fix_resnet.py:82-84)"""
        r = torch.round(x * float(2 ** self.fi))
        r = r.clamp_(-127.0, 127.0) if self.spec.sym else r.clamp_(0.0, 255.0)
        return r.int().float()

    def __call__(self, x_int):
        s = self.spec
        if s.kind == "fc":
            return F.linear(x_int, self.w, self.b)
        return F.conv2d(x_int, self.w, self.b, stride=s.stride, padding=s.pad, groups=s.groups)

    @property
    def fa(self):
        return self.fw + self.fi


def residual_add(res, res_fl, x, x_fl):
    fl = max(res_fl, x_fl)                                   # fix_resnet.py:91-98
    res = res * float(2 ** fl)
    res += x * float(2 ** fl)
    res = torch.clamp(res, max=float((1 << 31) - 1), min=float(-(1 << 31) + 1))
    return res / float(2 ** fl), fl


class IntegizeNet:
    """Float-dtype-holding-ints forward of one network, built from the IntModel state_dict."""

    def __init__(self, arch, sd, head_signed, dev):
        self.net = graph_for(arch, head_signed)
        self.L = {c.prefix: Layer(sd, c, dev) for c in self.net.convs()}

    def block(self, b, x, x_fl):
        res = x
        for j, c in enumerate(b.body):
            lay = self.L[c.prefix]
            res = lay(lay.quant_in(res))
            res.div_(float(2 ** lay.fa))
            if j + 1 < len(b.body) or b.relu_after_last:
                res = torch.relu_(res)
        res_fl = self.L[b.body[-1].prefix].fa
        if b.identity:
            res, res_fl = residual_add(res, res_fl, x, x_fl)
        elif b.shortcut is not None:
            sc = self.L[b.shortcut.prefix]
            s = sc(sc.quant_in(x))
            s.div_(float(2 ** sc.fa))
            res, res_fl = residual_add(res, res_fl, s, sc.fa)
        if b.post_relu:
            res = torch.relu_(res)
        return res, res_fl

    def forward(self, x_int):
        """x_int: float32 NCHW holding the head's 8-bit integers (what (x * 2^fi).int().float()
        yields, fix_resnet.py:385-390)."""
        net, head = self.net, self.L["head.0"]
        x = torch.relu_(head(x_int))
        if net.maxpool:
            x = F.max_pool2d(x, 3, 2, 1)
        fl = head.fa
        x.div_(float(2 ** fl))
        for b in net.blocks:
            x, fl = self.block(b, x, fl)
        if net.tail is not None:
            t = self.L[net.tail.prefix]
            x = torch.relu_(t(t.quant_in(x)).div_(float(2 ** t.fa)))
        x = x.sum(-1).sum(-1).div_(64.0)                     # FXQAvgPool2d float branch (:134-136)
        fc = self.L[net.fc.prefix]
        return fc(fc.quant_in(x))


def time_ms(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", 0)
    hs = synth.HEAD_SIGNED[a.arch]
    sd = synth.make_state_dict(a.arch, hs)
    x = synth.make_input(a.arch, a.batch, hs)
    xi = torch.from_numpy(x).to(dev)
    xf = xi.float()
    ref = IntegizeNet(a.arch, sd, hs, dev)
    eng = f8net_b200.compile(sd, arch=a.arch, head_signed=hs, chunk=a.batch)
    with torch.no_grad():
        y_int = eng(xi)
        y_flt = ref.forward(xf.clone())
        for _ in range(3):
            ref.forward(xf.clone())
        eager = time_ms(lambda: ref.forward(xf.clone()), a.iters)
        g = torch.cuda.CUDAGraph()
        xs = xf.clone()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ref.forward(xs.clone())
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            ref.forward(xs.clone())
        graph = time_ms(g.replay, a.iters)
        for _ in range(3):
            eng(xi)
        ours = time_ms(lambda: eng(xi), a.iters)
    diff = (y_int != y_flt)
    out = {"arch": a.arch, "batch": a.batch,
           "integize_restatement": {"ms_eager": eager, "ms_cuda_graph": graph,
                                    "images_per_s": a.batch / (min(eager, graph) / 1e3),
                                    "what": "plain-PyTorch restatement of IntModel's integize branch: float32 "
                                            "tensors holding the ints through cuDNN / cuBLAS, TF32 off"},
           "f8net_b200": {"ms": ours, "images_per_s": a.batch / (ours / 1e3),
                          "input": "int32 NCHW on the device (the reference's tensor)"},
           "speedup": min(eager, graph) / ours,
           "logits_differing_from_exact_int_path": int(diff.sum()), "logits_total": int(diff.numel()),
           "max_abs_logit_difference": float((y_int - y_flt).abs().max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
