"""Oracle forward passes of the reference's IntModel / IntBlock (int_op_only branch).

TEST INFRASTRUCTURE ONLY (see oracle/f8_oracle.c header).  Each function is a restatement
of one reference forward method, structured the same way so the two can be read side by
side; arithmetic primitives come from oracle.oracle (C).

A "state dict" here is a mapping  name -> numpy int32 array  with exactly the keys of the
reference's IntModel.state_dict():
    <p>.weight [O,C/g,kh,kw] | [O,K],  <p>.bias [O],  <p>.weight_fraclen (),  <p>.input_fraclen (1,)

Tensors carry their fraclen the way the reference does with a Python attribute: we use a
small (array, fraclen) tuple instead.
"""
import numpy as np

from . import oracle as O

# ----------------------------------------------------------------------------------------
# Topology tables, restated from the reference Model constructors.
# ----------------------------------------------------------------------------------------
# fix_resnet.py:447-457 block_setting_dict, feats :458, BasicBlock :122-153 / Bottleneck :224-256
RESNET_BLOCKS = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3],
                 152: [3, 8, 36, 3]}
RESNET_FEATS = [64, 128, 256, 512]
# fix_mobilenet_v1.py:176-183  [c, n, s]
MBV1_SETTING = [[64, 1, 1], [128, 2, 2], [256, 2, 2], [512, 6, 2], [1024, 2, 2]]
# fix_mobilenet_v2.py:282-291  [t, c, n, s]
MBV2_SETTING = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1],
                [6, 160, 3, 2], [6, 320, 1, 1]]


class Layer:
    """One int nn.Conv2d / nn.Linear: tensors + the Python attrs the state_dict lacks."""

    def __init__(self, sd, prefix, stride=1, pad=0, groups=1, sym=False):
        self.prefix = prefix
        self.w = np.ascontiguousarray(sd[prefix + ".weight"], dtype=np.int32)
        self.b = np.ascontiguousarray(sd[prefix + ".bias"], dtype=np.int32)
        self.fw = int(np.asarray(sd[prefix + ".weight_fraclen"]).reshape(-1)[0])
        self.fi = int(np.asarray(sd[prefix + ".input_fraclen"]).reshape(-1)[0])
        self.stride, self.pad, self.groups, self.sym = stride, pad, groups, bool(sym)

    def conv(self, x):
        return O.conv2d(x, self.w, self.b, self.stride, self.pad, self.groups)


def _trace(trace, key, val):
    if trace is not None:
        trace[key] = val


# Fixture calibration hook (tests/golden/make_golden.py): when set, called as
# CALIB(layer, tensor, fa) -> fi right before a layer's input requantisation so the
# generator can pick a non-degenerate input_fraclen.  Never set during parity checks.
CALIB = None


def _rq(layer, x, fa, trace):
    """res = int_op_only_fix_quant(res, 8, layer.input_fraclen, res.output_fraclen,
    layer.input_symmetric) -- the call every reference forward makes before an int layer."""
    if CALIB is not None:
        layer.fi = int(CALIB(layer, x, fa))
    q = O.requant(x, layer.fi, fa, layer.sym)
    _trace(trace, layer.prefix + ":in8", q)
    return q


def _run_body(body, x, fa, trace):
    """The conv loop shared by every IntBlock.forward:
    fix_resnet.py:28-39, fix_mobilenet_v1.py:27-38, fix_mobilenet_v2.py:22-33.
    body is a list of Layer | 'relu'."""
    res = x
    for item in body:
        if item == "relu":
            res = O.relu(res)
        else:
            res = _rq(item, res, fa, trace)
            res = item.conv(res)
            fa = item.fw + item.fi
            _trace(trace, item.prefix + ":acc", res)
    return res, fa


def _pool_fc(x, fa, fc, trace):
    """avgpool + requant + classifier + .float(): fix_resnet.py:367-383 (quant_avgpool True),
    fix_mobilenet_v1.py:131-147, fix_mobilenet_v2.py:225-241."""
    assert x.shape[2] == 7 and x.shape[3] == 7
    fa = fa + 6  # FXQAvgPool2d(7).shiftnum = round(log2(49)) = 6, fix_quant_ops.py:121-122
    assert fa <= 32  # fix_quant_ops.py:129
    p = O.avgpool_sum(x)
    _trace(trace, "avgpool:sum", p)
    q = _rq(fc, p, fa, trace)
    yi, yf = O.linear(q, fc.w, fc.b)
    _trace(trace, fc.prefix + ":acc", yi)
    return yf


# ----------------------------------------------------------------------------------------
# ResNet  (fix_resnet.py)
# ----------------------------------------------------------------------------------------
def resnet_forward(sd, depth, x, head_signed=False, trace=None, quant_maxpool=False):
    """IntModel.forward int branch, fix_resnet.py:352-383; IntBlock.forward :24-77.
    x: int32 [N,3,224,224] already in the head's 8-bit range (the head does not requantise)."""
    bottleneck = depth >= 50
    expansion = 4 if bottleneck else 1
    head = Layer(sd, "head.0", stride=2, pad=3, sym=head_signed)
    h = O.relu(head.conv(x))                       # head[:-1](x)            :358
    _trace(trace, "head.0:acc", h)
    if quant_maxpool:
        h = O.maxpool_int(h, 3, 2, 1)              # FXQMaxPool2d            :356
    else:
        h = O.maxpool_float_rt(h, 3, 2, 1)         # head[-1](x.float()).int() :359
    fa = head.fw + head.fi                         # :360-362
    _trace(trace, "head:pool", h)
    channels = 64
    for idx, n in enumerate(RESNET_BLOCKS[depth]):
        outp = RESNET_FEATS[idx] * expansion
        for i in range(n):
            stride = 2 if (i == 0 and idx != 0) else 1      # fix_resnet.py:462-465
            pfx = f"stage_{idx}_layer_{i}"
            if bottleneck:                                   # Bottleneck.int_block :308-319
                body = [Layer(sd, pfx + ".body.0"), "relu",
                        Layer(sd, pfx + ".body.2", stride=stride, pad=1), "relu",
                        Layer(sd, pfx + ".body.4")]
            else:                                            # BasicBlock.int_block :207-221
                body = [Layer(sd, pfx + ".body.0", stride=stride, pad=1), "relu",
                        Layer(sd, pfx + ".body.2", stride=1, pad=1)]
            identity = stride == 1 and channels == outp      # residual_connection :143, :246
            res, fr = _run_body(body, h, fa, trace)
            if identity:
                s, fs = h, fa
            else:
                sc = Layer(sd, pfx + ".shortcut.0", stride=stride)
                s = _rq(sc, h, fa, trace)                    # :57-59
                s = sc.conv(s)
                fs = sc.fw + sc.fi
                _trace(trace, sc.prefix + ":acc", s)
            h, fa = O.residual_add(res, s, fr, fs)           # :40-76
            h = O.relu(h)                                    # post_relu :77
            _trace(trace, pfx + ":out", h)
            channels = outp
    fc = Layer(sd, "classifier.0")
    return _pool_fc(h, fa, fc, trace)


# ----------------------------------------------------------------------------------------
# MobileNet V1  (fix_mobilenet_v1.py)
# ----------------------------------------------------------------------------------------
def mobilenet_v1_forward(sd, x, head_signed=False, trace=None):
    """IntModel.forward int branch, fix_mobilenet_v1.py:120-147; IntBlock.forward :23-38."""
    head = Layer(sd, "head.0", stride=2, pad=1, sym=head_signed)
    h = O.relu(head.conv(x))                                 # self.head(x) :123
    fa = head.fw + head.fi
    _trace(trace, "head.0:acc", h)
    channels = 32
    for idx, (c, n, s) in enumerate(MBV1_SETTING):
        for i in range(n):
            stride = s if i == 0 else 1                      # fix_mobilenet_v1.py:205-208
            pfx = f"stage_{idx}_layer_{i}"
            body = [Layer(sd, pfx + ".body.0", stride=stride, pad=1, groups=channels), "relu",
                    Layer(sd, pfx + ".body.2"), "relu"]      # int_block :82-92
            h, fa = _run_body(body, h, fa, trace)
            _trace(trace, pfx + ":out", h)
            channels = c
    fc = Layer(sd, "classifier.0")
    return _pool_fc(h, fa, fc, trace)


# ----------------------------------------------------------------------------------------
# MobileNet V2  (fix_mobilenet_v2.py)
# ----------------------------------------------------------------------------------------
def mobilenet_v2_forward(sd, x, head_signed=False, trace=None):
    """IntModel.forward int branch, fix_mobilenet_v2.py:207-241; IntBlock.forward :18-48."""
    head = Layer(sd, "head.0", stride=2, pad=1, sym=head_signed)
    h = O.relu(head.conv(x))
    fa = head.fw + head.fi
    _trace(trace, "head.0:acc", h)
    channels = 32
    for idx, (t, c, n, s) in enumerate(MBV2_SETTING):
        for i in range(n):
            stride = s if i == 0 else 1
            # double_side of the block-entry conv: fix_mobilenet_v2.py:311-331
            entry_sym = (idx != 0) if i == 0 else True
            pfx = f"stage_{idx}_layer_{i}"
            exp = channels * t
            if t != 1:                                       # InvertedResidual :87-108, int_block :168-176
                body = [Layer(sd, pfx + ".body.0", sym=entry_sym), "relu",
                        Layer(sd, pfx + ".body.2", stride=stride, pad=1, groups=exp), "relu",
                        Layer(sd, pfx + ".body.4")]
            else:
                body = [Layer(sd, pfx + ".body.0", stride=stride, pad=1, groups=exp,
                              sym=entry_sym), "relu",
                        Layer(sd, pfx + ".body.2")]
            identity = stride == 1 and channels == c         # :125
            res, fr = _run_body(body, h, fa, trace)
            if identity:
                h, fa = O.residual_add(res, h, fr, fa)       # :34-48, no ReLU afterwards
            else:
                h, fa = res, fr
            _trace(trace, pfx + ":out", h)
            channels = c
    tail = Layer(sd, "tail.0", sym=True)                     # double_side=True :338-347
    h = _rq(tail, h, fa, trace)                              # :217-219
    h = O.relu(tail.conv(h))                                 # self.tail(x) :220
    fa = tail.fw + tail.fi
    _trace(trace, "tail.0:acc", h)
    fc = Layer(sd, "classifier.0")
    return _pool_fc(h, fa, fc, trace)


def forward(arch, sd, x, head_signed=False, trace=None, quant_maxpool=False):
    """arch in {'resnet18','resnet50','mobilenet_v1','mobilenet_v2'}; x int32 NCHW."""
    x = np.ascontiguousarray(x, dtype=np.int32)
    if arch.startswith("resnet"):
        return resnet_forward(sd, int(arch[6:]), x, head_signed, trace, quant_maxpool)
    if arch == "mobilenet_v1":
        return mobilenet_v1_forward(sd, x, head_signed, trace)
    if arch == "mobilenet_v2":
        return mobilenet_v2_forward(sd, x, head_signed, trace)
    raise ValueError(arch)
