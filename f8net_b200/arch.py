"""Integer-graph description of the four F8Net networks on the int_op_only path.

The graph is what the reference's ``Model.int_model()`` builds (fix_resnet.py:526-544,
fix_mobilenet_v1.py:262-281, fix_mobilenet_v2.py:405-423) expressed as plain data:
every int ``nn.Conv2d`` / ``nn.Linear`` with the attributes the state_dict does NOT carry
(stride, padding, groups, ``input_symmetric``) plus the block wiring (ReLU positions,
identity / shortcut residuals, max-pool, avg-pool).  It can be derived two ways:

* ``graph_for(arch, head_signed)`` -- from the architecture name alone (needed when only a
  state_dict is available);
* ``graph_from_module(int_model)`` -- by walking a live reference ``IntModel``.
"""
from dataclasses import dataclass, field
from typing import List, Optional

ARCHS = ("resnet18", "resnet50", "mobilenet_v1", "mobilenet_v2")

# fix_resnet.py:447-458
_RESNET_BLOCKS = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3],
                  152: [3, 8, 36, 3]}
_RESNET_FEATS = [64, 128, 256, 512]
# fix_mobilenet_v1.py:176-183  (c, n, s)
_MBV1 = [(64, 1, 1), (128, 2, 2), (256, 2, 2), (512, 6, 2), (1024, 2, 2)]
# fix_mobilenet_v2.py:282-291  (t, c, n, s)
_MBV2 = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1),
         (6, 160, 3, 2), (6, 320, 1, 1)]


@dataclass
class ConvSpec:
    """One int nn.Conv2d (or nn.Linear when kind == 'fc')."""
    prefix: str            # state_dict prefix, e.g. 'stage_0_layer_1.body.2'
    cin: int
    cout: int
    k: int = 1
    stride: int = 1
    pad: int = 0
    groups: int = 1
    sym: bool = False      # input_symmetric (fix_quant_ops.py:710)
    kind: str = "conv"     # 'conv' | 'fc'

    @property
    def depthwise(self):
        return self.groups > 1

    def weight_shape(self):
        if self.kind == "fc":
            return (self.cout, self.cin)
        return (self.cout, self.cin // self.groups, self.k, self.k)


@dataclass
class BlockSpec:
    """One IntBlock: body convs with a ReLU between consecutive convs."""
    name: str
    body: List[ConvSpec]
    relu_after_last: bool = False          # MBV1 body ends with ReLU (fix_mobilenet_v1.py:85-90)
    shortcut: Optional[ConvSpec] = None    # ResNet downsample 1x1 (fix_resnet.py:216-217)
    identity: bool = False                 # residual_connection with the block input
    post_relu: bool = False                # ResNet post_relu after the add (fix_resnet.py:77)


@dataclass
class NetSpec:
    arch: str
    family: str                            # 'resnet' | 'mobilenet_v1' | 'mobilenet_v2'
    head: ConvSpec
    maxpool: bool                          # ResNet head max-pool 3x3 s2 p1 (fix_resnet.py:439)
    blocks: List[BlockSpec]
    tail: Optional[ConvSpec]               # MBV2 tail 1x1 + ReLU (fix_mobilenet_v2.py:338-349)
    fc: ConvSpec
    image_size: int = 224
    num_classes: int = 1000
    block_setting: list = field(default_factory=list)
    maxpool_int: bool = False              # FLAGS.quant_maxpool: FXQMaxPool2d, integer max without the
                                           # float round trip (fix_quant_ops.py:141-157, fix_resnet.py:355-356)

    def convs(self):
        """Every int layer in state_dict order."""
        out = [self.head]
        for b in self.blocks:
            out.extend(b.body)
            if b.shortcut is not None:
                out.append(b.shortcut)
        if self.tail is not None:
            out.append(self.tail)
        out.append(self.fc)
        return out


def _resnet(depth, head_signed, num_classes):
    bottleneck = depth >= 50
    exp = 4 if bottleneck else 1
    head = ConvSpec("head.0", 3, 64, 7, 2, 3, sym=head_signed)
    blocks, ch = [], 64
    for idx, n in enumerate(_RESNET_BLOCKS[depth]):
        outp = _RESNET_FEATS[idx] * exp
        for i in range(n):
            st = 2 if (i == 0 and idx != 0) else 1
            p = f"stage_{idx}_layer_{i}"
            if bottleneck:
                mid = outp // 4
                body = [ConvSpec(p + ".body.0", ch, mid, 1, 1, 0),
                        ConvSpec(p + ".body.2", mid, mid, 3, st, 1),
                        ConvSpec(p + ".body.4", mid, outp, 1, 1, 0)]
            else:
                body = [ConvSpec(p + ".body.0", ch, outp, 3, st, 1),
                        ConvSpec(p + ".body.2", outp, outp, 3, 1, 1)]
            ident = st == 1 and ch == outp
            sc = None if ident else ConvSpec(p + ".shortcut.0", ch, outp, 1, st, 0)
            blocks.append(BlockSpec(p, body, False, sc, ident, True))
            ch = outp
    fc = ConvSpec("classifier.0", ch, num_classes, kind="fc")
    return NetSpec(f"resnet{depth}", "resnet", head, True, blocks, None, fc,
                   block_setting=_RESNET_BLOCKS[depth])


def _mbv1(head_signed, num_classes):
    head = ConvSpec("head.0", 3, 32, 3, 2, 1, sym=head_signed)
    blocks, ch = [], 32
    for idx, (c, n, s) in enumerate(_MBV1):
        for i in range(n):
            st = s if i == 0 else 1
            p = f"stage_{idx}_layer_{i}"
            body = [ConvSpec(p + ".body.0", ch, ch, 3, st, 1, groups=ch),
                    ConvSpec(p + ".body.2", ch, c, 1, 1, 0)]
            blocks.append(BlockSpec(p, body, relu_after_last=True))
            ch = c
    fc = ConvSpec("classifier.0", ch, num_classes, kind="fc")
    return NetSpec("mobilenet_v1", "mobilenet_v1", head, False, blocks, None, fc,
                   block_setting=[list(t) for t in _MBV1])


def _mbv2(head_signed, num_classes):
    head = ConvSpec("head.0", 3, 32, 3, 2, 1, sym=head_signed)
    blocks, ch = [], 32
    for idx, (t, c, n, s) in enumerate(_MBV2):
        for i in range(n):
            st = s if i == 0 else 1
            entry_sym = (idx != 0) if i == 0 else True   # fix_mobilenet_v2.py:311-331
            p = f"stage_{idx}_layer_{i}"
            e = ch * t
            if t != 1:
                body = [ConvSpec(p + ".body.0", ch, e, 1, 1, 0, sym=entry_sym),
                        ConvSpec(p + ".body.2", e, e, 3, st, 1, groups=e),
                        ConvSpec(p + ".body.4", e, c, 1, 1, 0)]
            else:
                body = [ConvSpec(p + ".body.0", e, e, 3, st, 1, groups=e, sym=entry_sym),
                        ConvSpec(p + ".body.2", e, c, 1, 1, 0)]
            blocks.append(BlockSpec(p, body, False, None, st == 1 and ch == c, False))
            ch = c
    tail = ConvSpec("tail.0", ch, 1280, 1, 1, 0, sym=True)
    fc = ConvSpec("classifier.0", 1280, num_classes, kind="fc")
    return NetSpec("mobilenet_v2", "mobilenet_v2", head, False, blocks, tail, fc,
                   block_setting=[list(t) for t in _MBV2])


def graph_for(arch, head_signed=False, num_classes=1000, quant_maxpool=False):
    """NetSpec from the architecture name.  ``head_signed`` mirrors FLAGS.normalize
    (double_side of the head conv, fix_resnet.py:437-438); ``quant_maxpool`` mirrors
    FLAGS.quant_maxpool (ResNets: FXQMaxPool2d instead of nn.MaxPool2d on floats,
    fix_resnet.py:331-334).  FLAGS.quant_avgpool is always True on this path: the shipped
    int_op_only configs set it, and without it the reference leaves integer arithmetic
    (AdaptiveAvgPool2d on x.float(), fix_resnet.py:375-382)."""
    if arch.startswith("resnet"):
        net = _resnet(int(arch[6:]), head_signed, num_classes)
        net.maxpool_int = bool(quant_maxpool)
        return net
    if arch == "mobilenet_v1":
        return _mbv1(head_signed, num_classes)
    if arch == "mobilenet_v2":
        return _mbv2(head_signed, num_classes)
    raise ValueError(f"unknown arch {arch!r}; expected one of {ARCHS} or resnet<depth>")


# ----------------------------------------------------------------------------------------
# Deriving the graph from a live reference IntModel (duck-typed; torch only for isinstance)
# ----------------------------------------------------------------------------------------
def _conv_spec_from_module(m, prefix):
    import torch.nn as nn
    if isinstance(m, nn.Linear):
        return ConvSpec(prefix, m.in_features, m.out_features, kind="fc",
                        sym=bool(getattr(m, "input_symmetric", False)))
    assert isinstance(m, nn.Conv2d), type(m)
    k, st, pd = m.kernel_size, m.stride, m.padding
    if k[0] != k[1] or st[0] != st[1] or pd[0] != pd[1] or tuple(m.dilation) != (1, 1):
        raise ValueError(f"{prefix}: only square, undilated convolutions are on the F8Net path")
    return ConvSpec(prefix, m.in_channels, m.out_channels, k[0], st[0], pd[0], m.groups,
                    bool(getattr(m, "input_symmetric", False)))


def graph_from_module(im):
    """Walk a reference IntModel (any of the three model files) into a NetSpec."""
    import torch.nn as nn
    head = _conv_spec_from_module(im.head[0], "head.0")
    pools = [m for m in im.head if isinstance(m, nn.MaxPool2d)]
    maxpool = bool(pools)
    # FXQMaxPool2d subclasses nn.MaxPool2d (fix_quant_ops.py:141): told apart by class name
    maxpool_int = any(type(m).__name__ == "FXQMaxPool2d" for m in pools)
    for m in pools:
        k, s, p = (v if isinstance(v, int) else v[0] for v in (m.kernel_size, m.stride, m.padding))
        if (k, s, p) != (3, 2, 1):
            raise ValueError(f"head max-pool {k}x{k} s{s} p{p}: the F8Net head pools 3x3 s2 p1")
    # the integer tail is FXQAvgPool2d(7) (FLAGS.quant_avgpool); with nn.AdaptiveAvgPool2d the
    # reference averages in float32 (fix_resnet.py:375-382) -- different arithmetic, not this path
    if type(getattr(im, "avgpool", None)).__name__ != "FXQAvgPool2d":
        raise ValueError("IntModel.avgpool is not FXQAvgPool2d: build the model with quant_avgpool: True "
                         "(the int_op_only configs do); the float average pool is outside the integer path")
    has_tail = hasattr(im, "tail")
    blocks = []
    setting = list(im.block_setting)
    for idx, entry in enumerate(setting):
        if isinstance(entry, (list, tuple)):
            n = entry[1] if len(entry) == 3 else entry[2]
        else:
            n = entry
        for i in range(n):
            name = f"stage_{idx}_layer_{i}"
            blk = getattr(im, name)
            body, relu_last = [], False
            for j, layer in enumerate(blk.body):
                if isinstance(layer, nn.Conv2d):
                    body.append(_conv_spec_from_module(layer, f"{name}.body.{j}"))
                    relu_last = False
                elif isinstance(layer, nn.ReLU):
                    relu_last = True
                else:
                    raise ValueError(f"{name}.body.{j}: unexpected module {type(layer)}")
            sc = None
            if hasattr(blk, "shortcut"):
                sc = _conv_spec_from_module(blk.shortcut[0], f"{name}.shortcut.0")
            ident = bool(getattr(blk, "residual_connection", False))
            post_relu = hasattr(blk, "post_relu")
            blocks.append(BlockSpec(name, body, relu_last, sc, ident, post_relu))
    tail = _conv_spec_from_module(im.tail[0], "tail.0") if has_tail else None
    fc = _conv_spec_from_module(im.classifier[0], "classifier.0")
    if has_tail:
        family, arch = "mobilenet_v2", "mobilenet_v2"
    elif maxpool:
        family = "resnet"
        arch = "resnet?"
        for d, bs in _RESNET_BLOCKS.items():
            bott = len(blocks[0].body) == 3
            if bs == setting and (d >= 50) == bott:
                arch = f"resnet{d}"
    else:
        family, arch = "mobilenet_v1", "mobilenet_v1"
    return NetSpec(arch, family, head, maxpool, blocks, tail, fc,
                   num_classes=fc.cout, block_setting=setting, maxpool_int=maxpool_int)
