"""ONNX export of the integer graph.

Replaces ``onnx_export(model, data_shape, torch.int, device, file)`` for the int_op_only model
(/root/reference/myutils/export.py:4-31, called at /root/reference/fix_train.py:948-954): an
opset-11 ModelProto with one int32 input ``input`` [batch_size, 3, 224, 224], one float output
``output`` [batch_size, 1000], dynamic batch axis, and the network's parameters as int32
initialisers -- ``Conv`` / ``Gemm`` nodes carry the int32 weights and biases exactly as the
reference's tracer emits them for its int32 ``nn.Conv2d`` / ``nn.Linear``; the fixed-point
requantiser ``int_op_only_fix_quant`` (fix_quant_ops.py:90-114) is the subgraph

    r = Add(x, 2^(n-1));  tie = Equal(Mod(x, 2^n), 2^(n-1))
    q = Where(tie, Mul(FloorShift(r, n+1), 2), FloorShift(r, n));  y = Min(Max(q, lo), hi)
    FloorShift(v, k) = Div(Sub(v, Mod(v, 2^k)), 2^k)        (exact arithmetic shift: v - v mod 2^k
                                                              is divisible, so integer Div is exact)

(``Mul(x, 2^-n)`` for left shifts), the residual add is ``Mul`` / ``Add`` / ``Max(., INT_MIN+1)``,
the head max-pool is ``Cast(float) -> MaxPool -> Cast(int32)`` (or an integer ``MaxPool`` for
FXQMaxPool2d), FXQAvgPool2d is ``ReduceSum`` over H, W, and the logits are ``Cast`` to float.
All integer tensors are int32 with two's-complement wrap, as in the reference.

The ``onnx`` package is not needed: the ModelProto is written with a ~60-line protobuf wire
encoder (``_Msg``) following onnx.proto3 field numbers; ``read_model`` decodes it again (tests
re-execute the decoded graph and compare with the engine's logits).
"""
import struct
from typing import Dict, List

import numpy as np

from .arch import NetSpec, graph_for

# onnx.proto3 TensorProto.DataType
FLOAT, INT32, INT64, BOOL = 1, 6, 7, 9
# AttributeProto.AttributeType
A_FLOAT, A_INT, A_STRING, A_INTS = 1, 2, 3, 7
OPSET = 11
IR_VERSION = 6          # the IR version onnx 1.6 / opset 11 writers emit


# ----------------------------------------------------------------------------------------------
# protobuf wire format
# ----------------------------------------------------------------------------------------------
def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


class _Msg:
    def __init__(self):
        self.b = bytearray()

    def varint(self, field, v):
        self.b += _varint(field << 3) + _varint(int(v))
        return self

    def bytes_(self, field, data):
        if isinstance(data, _Msg):
            data = bytes(data.b)
        elif isinstance(data, str):
            data = data.encode()
        self.b += _varint((field << 3) | 2) + _varint(len(data)) + data
        return self

    def float_(self, field, v):
        self.b += _varint((field << 3) | 5) + struct.pack("<f", v)
        return self


def _tensor(name, arr):
    arr = np.ascontiguousarray(arr)
    dt = {np.dtype(np.int32): INT32, np.dtype(np.int64): INT64, np.dtype(np.float32): FLOAT}[arr.dtype]
    t = _Msg()
    for d in arr.shape:
        t.varint(1, d)                       # dims
    t.varint(2, dt)                          # data_type
    t.bytes_(8, name)                        # name
    t.bytes_(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())     # raw_data
    return t


def _attr(name, value):
    a = _Msg().bytes_(1, name)
    if isinstance(value, (list, tuple)):
        for v in value:
            a.varint(8, v)                   # ints
        a.varint(20, A_INTS)
    elif isinstance(value, float):
        a.float_(2, value).varint(20, A_FLOAT)
    elif isinstance(value, str):
        a.bytes_(4, value).varint(20, A_STRING)
    else:
        a.varint(3, value).varint(20, A_INT)
    return a


def _value_info(name, elem_type, dims):
    shape = _Msg()
    for d in dims:
        dim = _Msg()
        if isinstance(d, str):
            dim.bytes_(2, d)                 # dim_param
        else:
            dim.varint(1, d)                 # dim_value
        shape.bytes_(1, dim)
    tt = _Msg().varint(1, elem_type).bytes_(2, shape)
    return _Msg().bytes_(1, name).bytes_(2, _Msg().bytes_(1, tt))


class GraphBuilder:
    """Accumulates NodeProtos and initialisers; tensor names are generated."""

    def __init__(self):
        self.nodes: List[_Msg] = []
        self.inits: List[_Msg] = []
        self._n = 0
        self._consts: Dict[tuple, str] = {}

    def fresh(self, hint):
        self._n += 1
        return f"{hint}_{self._n}"

    def init(self, name, arr):
        self.inits.append(_tensor(name, arr))
        return name

    def const(self, value, dtype=np.int32):
        key = (int(value), np.dtype(dtype).str)
        if key not in self._consts:
            self._consts[key] = self.init(f"c{'i' if dtype == np.int32 else 'l'}_{value}".replace("-", "m"),
                                          np.array(value, dtype=dtype))
        return self._consts[key]

    def node(self, op, inputs, hint=None, **attrs):
        out = self.fresh(hint or op.lower())
        n = _Msg()
        for i in inputs:
            n.bytes_(1, i)
        n.bytes_(2, out).bytes_(3, out).bytes_(4, op)
        for k, v in attrs.items():
            n.bytes_(5, _attr(k, v))
        self.nodes.append(n)
        return out


def _wrap_i32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def build_graph(net: NetSpec, sd) -> _Msg:
    """GraphProto of IntModel.forward's int_op_only branch for ``net`` with the parameters ``sd``
    (reference state_dict layout: int32 weights / biases, weight_fraclen, input_fraclen)."""
    g = GraphBuilder()

    def fl(prefix):
        fw = int(np.asarray(sd[prefix + ".weight_fraclen"]).reshape(-1)[0])
        fi = int(np.asarray(sd[prefix + ".input_fraclen"]).reshape(-1)[0])
        return fw, fi

    def floor_shift(v, k):
        p = g.const(1 << k)
        return g.node("Div", [g.node("Sub", [v, g.node("Mod", [v, p])]), p], "sra")

    def requant(x, layer, fa):
        # int_op_only_fix_quant(x, 8, fi, fa, sym), fix_quant_ops.py:90-114
        _, fi = fl(layer.prefix)
        n = fa - fi
        if n > 0:
            half = g.const(1 << (n - 1))
            r = g.node("Add", [x, half])
            tie = g.node("Equal", [g.node("Mod", [x, g.const(1 << n)]), half], "tie")
            q = g.node("Where", [tie, g.node("Mul", [floor_shift(r, n + 1), g.const(2)]), floor_shift(r, n)], "rhe")
        elif n < 0:
            q = g.node("Mul", [x, g.const(_wrap_i32(1 << (-n)))], "shl")
        else:
            q = x
        lo, hi = (-127, 127) if layer.sym else (0, 255)
        return g.node("Min", [g.node("Max", [q, g.const(lo)]), g.const(hi)], "q8")

    def conv(x, layer):
        w = g.init(layer.prefix + ".weight", np.asarray(sd[layer.prefix + ".weight"], dtype=np.int32))
        b = g.init(layer.prefix + ".bias", np.asarray(sd[layer.prefix + ".bias"], dtype=np.int32))
        if layer.kind == "fc":
            return g.node("Gemm", [x, w, b], "fc", alpha=1.0, beta=1.0, transB=1)
        return g.node("Conv", [x, w, b], "conv", dilations=[1, 1], group=layer.groups,
                      kernel_shape=[layer.k, layer.k], pads=[layer.pad] * 4, strides=[layer.stride] * 2)

    def body(x, fa, layers, relu_after_last):
        r, fr = x, fa
        for i, L in enumerate(layers):
            r = conv(requant(r, L, fr), L)
            fr = sum(fl(L.prefix))
            if i < len(layers) - 1 or relu_after_last:
                r = g.node("Relu", [r])
        return r, fr

    # ---- head (no requant of the input: fix_resnet.py:355-358) ----
    x = conv("input", net.head)
    x = g.node("Relu", [x])
    fa = sum(fl(net.head.prefix))
    if net.maxpool:
        if net.maxpool_int:      # FXQMaxPool2d: pad with zeros, integer max (fix_quant_ops.py:141-157)
            x = g.node("MaxPool", [x], "maxpool", kernel_shape=[3, 3], pads=[1, 1, 1, 1], strides=[2, 2])
        else:                    # self.head[-1](x.float()).int()  (fix_resnet.py:358-359)
            f = g.node("Cast", [x], "tofloat", to=FLOAT)
            f = g.node("MaxPool", [f], "maxpool", kernel_shape=[3, 3], pads=[1, 1, 1, 1], strides=[2, 2])
            x = g.node("Cast", [f], "toint", to=INT32)
    # ---- blocks ----
    for blk in net.blocks:
        r, fr = body(x, fa, blk.body, blk.relu_after_last)
        if blk.identity or blk.shortcut is not None:
            if blk.identity:
                s, fs = x, fa
            else:
                s = conv(requant(x, blk.shortcut, fa), blk.shortcut)
                fs = sum(fl(blk.shortcut.prefix))
            # align fraclens by a wrapping left shift, add, clamp_(min=INT_MIN+1) (fix_resnet.py:62-76)
            if fr > fs:
                s = g.node("Mul", [s, g.const(_wrap_i32(1 << (fr - fs)))], "align")
            elif fs > fr:
                r = g.node("Mul", [r, g.const(_wrap_i32(1 << (fs - fr)))], "align")
            r = g.node("Max", [g.node("Add", [r, s], "residual"), g.const(-(1 << 31) + 1)], "clamp")
            fr = max(fr, fs)
            if blk.post_relu:
                r = g.node("Relu", [r])
        x, fa = r, fr
    if net.tail is not None:     # fix_mobilenet_v2.py:217-220
        x = g.node("Relu", [conv(requant(x, net.tail, fa), net.tail)])
        fa = sum(fl(net.tail.prefix))
    # ---- FXQAvgPool2d (int64 sum -> int32 wrap) + requant + classifier + .float() ----
    s64 = g.node("ReduceSum", [g.node("Cast", [x], "tolong", to=INT64)], "avgpool", axes=[2, 3], keepdims=0)
    x = g.node("Cast", [s64], "wrap", to=INT32)
    fa += 6
    x = conv(requant(x, net.fc, fa), net.fc)
    out = g.node("Cast", [x], "logits", to=FLOAT)
    g.nodes.append(_Msg().bytes_(1, out).bytes_(2, "output").bytes_(3, "output_identity").bytes_(4, "Identity"))

    gp = _Msg()
    for n in g.nodes:
        gp.bytes_(1, n)
    gp.bytes_(2, f"f8net_int_op_only_{net.arch}")
    for t in g.inits:
        gp.bytes_(5, t)
    S = net.image_size
    gp.bytes_(11, _value_info("input", INT32, ["batch_size", 3, S, S]))
    gp.bytes_(12, _value_info("output", FLOAT, ["batch_size", net.num_classes]))
    return gp


def export_onnx(state_dict, path, arch=None, head_signed=False, quant_maxpool=False, net: NetSpec = None):
    """Write the int_op_only model as an opset-11 ONNX file; returns the number of bytes written."""
    from .engine import _to_numpy_sd, infer_arch
    sd = state_dict() if callable(state_dict) else state_dict
    sd = _to_numpy_sd(sd)
    if net is None:
        net = graph_for(arch or infer_arch(sd), bool(head_signed), quant_maxpool=bool(quant_maxpool))
    m = _Msg()
    m.varint(1, IR_VERSION)
    m.bytes_(2, "f8net_b200")                                   # producer_name
    m.bytes_(3, "2")                                            # producer_version
    m.bytes_(7, build_graph(net, sd))
    m.bytes_(8, _Msg().bytes_(1, "").varint(2, OPSET))          # opset_import
    data = bytes(m.b)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


# ----------------------------------------------------------------------------------------------
# reader (round-trip tests; also lets a user inspect the file without the onnx package)
# ----------------------------------------------------------------------------------------------
def _fields(buf):
    i, n = 0, len(buf)
    while i < n:
        key = shift = 0
        while True:
            b = buf[i]
            i += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        field, wt = key >> 3, key & 7
        if wt == 0:
            v = shift = 0
            while True:
                b = buf[i]
                i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield field, v
        elif wt == 2:
            ln = shift = 0
            while True:
                b = buf[i]
                i += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield field, bytes(buf[i:i + ln])
            i += ln
        elif wt == 5:
            yield field, struct.unpack("<f", buf[i:i + 4])[0]
            i += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def read_model(path):
    """Decode a file written by ``export_onnx``: dict(ir_version, opset, graph_name, inputs, outputs,
    initializers {name: ndarray}, nodes [dict(op, inputs, outputs, attrs)])."""
    data = open(path, "rb").read()
    out = {"nodes": [], "initializers": {}, "inputs": [], "outputs": []}
    graph = None
    for f, v in _fields(data):
        if f == 1:
            out["ir_version"] = v
        elif f == 7:
            graph = v
        elif f == 8:
            out["opset"] = dict(_fields(v)).get(2)
    for f, v in _fields(graph):
        if f == 1:
            node = {"inputs": [], "outputs": [], "attrs": {}}
            for nf, nv in _fields(v):
                if nf == 1:
                    node["inputs"].append(nv.decode())
                elif nf == 2:
                    node["outputs"].append(nv.decode())
                elif nf == 4:
                    node["op"] = nv.decode()
                elif nf == 5:
                    name, ints, val = None, [], None
                    for af, av in _fields(nv):
                        if af == 1:
                            name = av.decode()
                        elif af == 8:
                            ints.append(_signed(av))
                        elif af == 3:
                            val = _signed(av)
                        elif af == 2:
                            val = av
                        elif af == 4:
                            val = av.decode()
                        elif af == 20:
                            kind = av
                    node["attrs"][name] = ints if kind == A_INTS else val
            out["nodes"].append(node)
        elif f == 2:
            out["graph_name"] = v.decode()
        elif f == 5:
            dims, dt, name, raw = [], None, None, b""
            for tf, tv in _fields(v):
                if tf == 1:
                    dims.append(tv)
                elif tf == 2:
                    dt = tv
                elif tf == 8:
                    name = tv.decode()
                elif tf == 9:
                    raw = tv
            np_dt = {INT32: "<i4", INT64: "<i8", FLOAT: "<f4"}[dt]
            out["initializers"][name] = np.frombuffer(raw, dtype=np_dt).reshape(dims).copy()
        elif f in (11, 12):
            info = {"dims": []}
            for vf, vv in _fields(v):
                if vf == 1:
                    info["name"] = vv.decode()
                elif vf == 2:
                    tt = dict(_fields(dict(_fields(vv))[1]))
                    info["elem_type"] = tt[1]
                    for sf, sv in _fields(tt[2]):
                        d = dict(_fields(sv))
                        info["dims"].append(d[2].decode() if 2 in d else d.get(1, 0))
            out["inputs" if f == 11 else "outputs"].append(info)
    return out
