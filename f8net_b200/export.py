"""Offline export: a float-simulation F8Net checkpoint -> the int32 state_dict of ``IntModel``.

SURVEY.md 8(f) rank 2 / row A5.  The reference converts a trained float-sim ``Model`` with
``Model.int_model()`` (fix_resnet.py:526-544, fix_mobilenet_v1.py:262-281,
fix_mobilenet_v2.py:405-423), which calls per layer ``int_conv()`` / ``int_fc()``
(fix_quant_ops.py:680-714, :1165-1195) -> ``int_weight`` / ``int_bias`` /
``get_weight_fraclen`` (:583-615, :661-678, :1063-1096, :1147-1163) on top of
``float_weight`` / ``float_bias`` / ``fix_scaling`` (:486-505, :533-581, :1022-1061).  That needs
the reference's module tree; this module needs only the checkpoint's ``state_dict`` (the
``best_model.pt`` the reference trains and loads at fix_train.py:877-891) and the handful of
config flags the conversion reads.  The result has exactly the keys, shapes and int32 dtypes of
the reference ``IntModel.state_dict()`` and feeds ``f8net_b200.compile``.

Not on the per-image path (once per model), so it is host code like the reference's: the float
arithmetic is done with torch CPU float32 tensor ops written in the reference's own operation
order, which is what makes the integers bit-identical (tests/golden/make_export_golden.py pins
that against the unmodified reference for all four networks).

Layer wiring restated here (the float-sim modules keep it in Python attributes, not in the
checkpoint):
  * ``following`` layer -- whose ``fix_scaling`` divides this layer's folded weight / bias:
    head -> first block's body[0]; body[i] -> body[i+1]; a block's last conv and its shortcut ->
    the next block's body[0] (fix_resnet.py:194-199, 294-300, 466-467, 484;
    fix_mobilenet_v1.py:71-74, 213, 230; fix_mobilenet_v2.py:156-160, 335, 350, 370);
  * ``master`` layer -- whose ``alpha`` (and ``input_fraclen`` when sharing) a block-entry conv and
    the shortcut reuse: the body[0] of the previous block if that block had an identity residual,
    recursively (fix_resnet.py:148-153, 456-466; fix_mobilenet_v2.py:126-131, 313-334, 347).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from .arch import NetSpec, graph_for


@dataclass
class ExportFlags:
    """The FLAGS the conversion reads (defaults = the *_int_op_only*.yml configs)."""
    weight_format: tuple = (8, 7)
    input_format: tuple = (8, 6)
    format_from_metric: bool = True
    format_grid_search: bool = False      # res50 tiny_finetuning config: True
    metric: str = "std"
    no_clipping: bool = False             # res50 tiny_finetuning config: True
    input_fraclen_sharing: bool = False
    rescale_forward: bool = True          # ReLUClipFXQLinear only (fix_train.py:293-294)
    rescale_forward_conv: bool = False
    rescale_type: str = "constant"
    quant_avgpool: bool = True
    normalize: bool = False               # head double_side / weight_only (fix_resnet.py:437-438)
    bn_eps: float = 1e-5
    avgpool_kernel: int = 7


@dataclass
class _Layer:
    fprefix: str                 # float-sim module prefix, e.g. 'stage_0_layer_0.body.1'
    iprefix: str                 # IntModel prefix, e.g. 'stage_0_layer_0.body.2'
    kind: str                    # 'conv' | 'fc'
    groups: int = 1
    out_ch: int = 0
    k: int = 1
    weight_only: bool = False
    double_side: bool = False
    bita_min: Optional[int] = None
    master: Optional["_Layer"] = None
    following: Optional["_Layer"] = None
    avgpool_scale: float = 1.0
    extra: dict = field(default_factory=dict)


def float_layers(net: NetSpec, flags: ExportFlags) -> List[_Layer]:
    """The float-sim layer list of ``Model.__init__`` with master / following wiring."""
    ds_head = bool(flags.normalize)
    head = _Layer("head.0", "head.0", "conv", 1, net.head.cout, net.head.k,
                  weight_only=not ds_head, double_side=ds_head, bita_min=8)
    layers = [head]
    prev_last: List[_Layer] = [head]            # layers whose `following` is the next master_child
    master = None
    for b in net.blocks:
        body = []
        for j, c in enumerate(b.body):
            # float body has no ReLU modules between convs: int body.{0,2,4} <-> float body.{0,1,2}
            body.append(_Layer(f"{b.name}.body.{j}", c.prefix, "conv", c.groups, c.cout, c.k,
                               double_side=bool(c.sym)))
        body[0].master = master if net.family != "mobilenet_v1" else None
        sc = None
        if b.shortcut is not None:
            sc = _Layer(f"{b.name}.shortcut.0", b.shortcut.prefix, "conv", 1, b.shortcut.cout, 1,
                        double_side=bool(b.shortcut.sym), master=master)
        for p in prev_last:
            p.following = body[0]
        for j in range(len(body) - 1):
            body[j].following = body[j + 1]
        prev_last = [body[-1]] + ([sc] if sc is not None else [])
        layers.extend(body)
        if sc is not None:
            layers.append(sc)
        if net.family != "mobilenet_v1":
            master = body[0] if b.identity else None
    last_conv = layers[-1] if net.blocks[-1].shortcut is None else layers[-2]
    if net.tail is not None:
        tail = _Layer("tail.0", "tail.0", "conv", 1, net.tail.cout, 1, double_side=bool(net.tail.sym),
                      master=master)
        for p in prev_last:
            p.following = tail
        prev_last = [tail]
        layers.append(tail)
        last_conv = tail
    fc = _Layer("classifier.0", "classifier.0", "fc", 1, net.fc.cout, 1, double_side=bool(net.fc.sym))
    for p in prev_last:
        p.following = fc
    layers.append(fc)
    if flags.quant_avgpool:
        # FXQAvgPool2d.scale = 2^round(log2 k^2) / k^2 (fix_quant_ops.py:121-123), applied to the last
        # conv before the pool by int_model() (fix_resnet.py:527-539; MBV2: the tail, :419)
        import math
        k2 = flags.avgpool_kernel ** 2
        last_conv.avgpool_scale = 2 ** int(round(math.log2(k2))) / k2
    return layers


class _Exporter:
    def __init__(self, sd, flags: ExportFlags):
        import torch
        self.t = torch
        self.sd = sd
        self.f = flags

    def p(self, layer: _Layer, name):
        return self.sd[f"{layer.fprefix}.{name}"].detach().to(self.t.float32).cpu()

    # ---- fix_quant_ops.py:64-87 ------------------------------------------------------------
    def fix_quant(self, x, wl, fl, align_dim, signed):
        t = self.t
        expand = x.dim() - align_dim - 1
        fl = fl[(...,) + (None,) * expand]
        res = x * (2 ** fl)
        res.round_()
        bound = 2 ** (wl - 1) - 1 if signed else 2 ** wl - 1
        res.clamp_(max=bound, min=-bound if signed else 0)
        res.div_(2 ** fl)
        return res

    # ---- fix_quant_ops.py:17-37 ------------------------------------------------------------
    def weight_fraclen(self, w, wl, align_dim, is_fc):
        t = self.t
        if self.f.format_grid_search:
            errs = []
            for fl in range(wl + 1 - 1):
                res = self.fix_quant(w, wl, t.ones(w.shape[align_dim]) * fl * 1.0, align_dim, True)
                errs.append(t.mean((w - res) ** 2) ** 0.5)
            return t.argmin(t.tensor(errs)) * 1.0
        if not self.f.format_from_metric:
            raise NotImplementedError("needs format_from_metric or format_grid_search")
        assert wl == 8, "Word length other than 8bit has not been implemented"
        axes = (0, 1) if is_fc else (0, 1, 2, 3)
        if self.f.metric == "std":
            m, coeff = t.std(w, axis=axes), 40
        elif self.f.metric == "mae":
            m, coeff = t.mean(t.abs(w), axis=axes), 30
        elif self.f.metric == "rms":
            m, coeff = t.mean(w ** 2, axis=axes) ** 0.5, 40
        else:
            raise NotImplementedError(self.f.metric)
        fl = t.floor(t.log2(coeff * 1 / m))
        fl.clamp_(max=8 - 1, min=0)
        return t.clamp(fl, max=wl - 1, min=0)

    # ---- get_alpha / get_input_fraclen / fix_scaling (fix_quant_ops.py:452-505) -------------
    def alpha(self, layer: _Layer):
        if layer.master is not None:
            return self.alpha(layer.master)
        a = self.p(layer, "alpha")
        return self.t.ones_like(a) if layer.weight_only else a

    def input_fraclen_raw(self, layer: _Layer):
        if layer.weight_only:
            return self.t.ones_like(self.p(layer, "input_fraclen")) * 8
        if layer.master is not None and self.f.input_fraclen_sharing:
            return self.input_fraclen_raw(layer.master)
        return self.p(layer, "input_fraclen")

    def x_wl(self, layer: _Layer):
        wl = self.f.input_format[0]
        return max(wl, layer.bita_min) if layer.bita_min is not None else wl

    def input_fraclen(self, layer: _Layer):
        fi = self.t.round(self.input_fraclen_raw(layer))
        return self.t.clamp(fi, max=self.x_wl(layer) - int(layer.double_side), min=0)

    def fix_scaling(self, layer: _Layer):
        alpha = self.t.abs(self.alpha(layer))
        if self.f.no_clipping:
            return self.t.ones_like(alpha)
        if layer.weight_only:
            return alpha
        return 2 ** self.input_fraclen(layer) * alpha / (2 ** (self.x_wl(layer) - int(layer.double_side)) - 1)

    # ---- conv: float_weight / float_bias / int_weight / int_bias (fix_quant_ops.py:533-615) --
    def conv(self, layer: _Layer):
        t = self.t
        weight = self.p(layer, "conv.weight")
        if self.f.rescale_forward_conv:
            if self.f.rescale_type == "stddev":
                ws = t.std(weight)
            elif self.f.rescale_type == "constant":
                # the reference cannot export this combination either: float_weight reads
                # self.out_channels / self.kernel_size, which ReLUClipFXQConvBN does not define
                # (fix_quant_ops.py:539-541 -> AttributeError inside the int_weight property)
                raise NotImplementedError("rescale_forward_conv with rescale_type 'constant' is not exportable "
                                          "in the reference (fix_quant_ops.py:539-541)")
            else:
                raise NotImplementedError
            ws = ws / t.std(weight)
        else:
            ws = 1.0
        weight = weight * ws
        bn_w, bn_b = self.p(layer, "bn.weight"), self.p(layer, "bn.bias")
        bn_mean = self.p(layer, "bn.running_mean")
        bn_std = t.sqrt(self.p(layer, "bn.running_var") + self.f.bn_eps)
        fs, fs_next = self.fix_scaling(layer), self.fix_scaling(layer.following)
        if layer.groups == 1:
            fw_ = (bn_w / bn_std)[:, None, None, None] * weight * fs[(...,) + (None, None)] / \
                fs_next[(...,) + (None, None, None)]
        else:
            fw_ = (bn_w / bn_std)[:, None, None, None] * weight * fs[(...,) + (None, None, None)] / \
                fs_next[(...,) + (None, None, None)]
        fb_ = (bn_b - bn_w / bn_std * bn_mean) / fs_next
        w_wl = self.f.weight_format[0]
        wfl = self.weight_fraclen(fw_ * layer.avgpool_scale, w_wl, 0, False)
        q = self.fix_quant(fw_ * layer.avgpool_scale, w_wl, wfl, 0, True)
        int_w = (q * (2 ** wfl)).int()
        fi = self.input_fraclen(layer)
        b = self.fix_quant(fb_ * layer.avgpool_scale, 32, fi + wfl, 0, True)
        int_b = (b * (2 ** (fi + wfl))).int()
        return int_w, int_b, wfl.int(), fi.int()

    # ---- linear (fix_quant_ops.py:1022-1096, 1147-1195) --------------------------------------
    def fc(self, layer: _Layer):
        t = self.t
        weight = self.p(layer, "weight")
        w_wl = self.f.weight_format[0]
        wfl = self.weight_fraclen(weight, w_wl, 1, True)
        int_w = (self.fix_quant(weight, w_wl, wfl, 0, True) * (2 ** wfl)).int()
        # float_bias re-derives the fraclen with align_dim 0 (same value) and quantises a copy
        wq = self.fix_quant(weight * 1.0, w_wl, self.weight_fraclen(weight * 1.0, w_wl, 0, True), 0, True)
        if self.f.rescale_forward:
            if self.f.rescale_type == "stddev":
                ws = t.std(weight)
            elif self.f.rescale_type == "constant":
                ws = 1.0 / (layer.out_ch) ** 0.5
            else:
                raise NotImplementedError
            ws = ws / t.std(wq)
        else:
            ws = 1.0
        fb_ = self.p(layer, "bias") / self.fix_scaling(layer) / ws
        fi = self.input_fraclen(layer)
        b = self.fix_quant(fb_, 32, fi + wfl, 0, True)
        int_b = (b * (2 ** (fi + wfl))).int()
        return int_w, int_b, wfl.int(), fi.int()


def export_int_state_dict(float_state_dict, arch: str, flags: Optional[ExportFlags] = None,
                          num_classes: int = 1000) -> Dict[str, "object"]:
    """``Model.int_model().state_dict()`` computed from the float-sim ``state_dict`` alone.

    ``float_state_dict``: the reference's training checkpoint (``best_model.pt`` -> its
    ``'model'`` entry; a ``module.`` DataParallel prefix is stripped).  Returns int32 torch
    tensors keyed ``<prefix>.{weight,bias,weight_fraclen,input_fraclen}`` in IntModel order."""
    import torch
    flags = flags or ExportFlags()
    sd = float_state_dict
    if isinstance(sd, dict) and "model" in sd and not any(k.endswith(".weight") for k in sd):
        sd = sd["model"]
    if callable(sd):
        sd = sd()
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    net = graph_for(arch, head_signed=bool(flags.normalize), num_classes=num_classes)
    layers = float_layers(net, flags)
    ex = _Exporter(sd, flags)
    by_int = {}
    with torch.no_grad():
        for layer in layers:
            need = "weight" if layer.kind == "fc" else "conv.weight"
            if f"{layer.fprefix}.{need}" not in sd:
                raise KeyError(f"{layer.fprefix}.{need} missing: not a float-sim {arch} checkpoint")
            by_int[layer.iprefix] = ex.fc(layer) if layer.kind == "fc" else ex.conv(layer)
    out = {}
    for c in net.convs():                     # IntModel.state_dict() key order
        w, b, wfl, fi = by_int[c.prefix]
        out[c.prefix + ".weight"] = w.contiguous()
        out[c.prefix + ".bias"] = b.contiguous()
        out[c.prefix + ".weight_fraclen"] = wfl
        out[c.prefix + ".input_fraclen"] = fi
    return out


def compile_float(float_state_dict, arch: str, flags: Optional[ExportFlags] = None, **kw):
    """Float-sim checkpoint -> Engine: export_int_state_dict + f8net_b200.compile."""
    from .engine import compile as _compile
    flags = flags or ExportFlags()
    sd = export_int_state_dict(float_state_dict, arch, flags)
    return _compile(sd, arch=arch, head_signed=bool(flags.normalize), **kw)
