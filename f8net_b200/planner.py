"""Host-side planner: reference integer graph (arch.NetSpec + IntModel.state_dict tensors)
-> the fused launch list and buffer table of include/f8b200.h (f8_op / f8_buffer).

The fusion rule: the consumer-side ``int_op_only_fix_quant`` of layer
L+1 (/root/reference/models/fix_quant_ops.py:90-114) runs in the epilogue of the kernel that
produces L+1's input -- all its parameters (input_fraclen, input_symmetric, the producer's
accumulator fraclen) are plan-time constants -- after the ReLU where the reference has one,
and the int32 value is additionally stored wherever a residual add or the average pool reads
it.  A tensor consumed by two int layers with different (input_fraclen, input_symmetric)
(downsample-block input: body[0] and shortcut[0], /root/reference/models/fix_resnet.py:28-59)
gets two 8-bit images.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _capi as C
from .arch import ConvSpec, NetSpec

CH_ALIGN = 16


def cpad(c):
    return (c + CH_ALIGN - 1) // CH_ALIGN * CH_ALIGN


@dataclass
class Buf:
    idx: int
    bytes_per_image: int
    first: int = -1          # op index that writes it
    last: int = -1           # last op index that reads it
    offset: int = 0
    name: str = ""


@dataclass
class Op:
    kind: int
    name: str
    cin: int = 0
    cout: int = 0
    cin_pad: int = 0
    cout_pad: int = 0
    k: int = 1
    stride: int = 1
    pad: int = 0
    hin: int = 1
    win: int = 1
    hout: int = 1
    wout: int = 1
    in_signed: int = 0
    in_buf: int = -1
    weight: Optional[np.ndarray] = None
    bias: Optional[np.ndarray] = None
    carry_in_buf: int = -1
    carry_shift: int = 0
    relu: int = 0
    carry_out_buf: int = -1
    outs: List[Tuple[int, int, int]] = field(default_factory=list)   # (buf, shift, signed)
    out_f32: int = 0
    flags: int = 0
    # planning-only
    fa: int = 0                      # fraclen of the int32 value this op produces
    index: int = -1


class Plan:
    """Ordered ops + buffers; ``to_desc()`` yields the ctypes descriptor."""

    def __init__(self, net: NetSpec):
        self.net = net
        self.ops: List[Op] = []
        self.bufs: List[Buf] = []
        self.workspace_per_image = 0

    # -- construction helpers -----------------------------------------------------------
    def new_buf(self, nbytes, name):
        b = Buf(len(self.bufs), int(nbytes), name=name)
        self.bufs.append(b)
        return b.idx

    def emit(self, op: Op):
        op.index = len(self.ops)
        self.ops.append(op)
        return op

    def image8(self, producer: Op, shift: int, signed: bool):
        """8-bit image of ``producer``'s int32 value requantised with (shift, signed);
        created on first request (at most two per producer)."""
        for b, s, g in producer.outs:
            if s == shift and g == int(signed):
                return b
        if len(producer.outs) == 2:
            raise ValueError(f"{producer.name}: more than two distinct requantised consumers")
        b = self.new_buf(producer.hout * producer.wout * producer.cout_pad,
                         f"{producer.name}:q{shift}{'s' if signed else 'u'}")
        producer.outs.append((b, shift, int(signed)))
        return b

    def carry(self, producer: Op):
        """int32 image of ``producer``'s value in the engine's pixel-interleaved carry layout
        (blocks of 128 pixels x 4 channels, csrc/f8_common.cuh): the pixel count is rounded up
        to 128 per image so that any batch fits."""
        if producer.carry_out_buf < 0:
            pixels = (producer.hout * producer.wout + 127) // 128 * 128
            producer.carry_out_buf = self.new_buf(pixels * producer.cout_pad * 4,
                                                  f"{producer.name}:i32")
        return producer.carry_out_buf

    # -- liveness + offsets -------------------------------------------------------------
    def finalize(self, keep_buffers=False):
        """Liveness + first-fit offsets.  ``keep_buffers``: every buffer gets its own range (no reuse),
        so that all intermediate tensors survive the run (per-layer parity, Engine.read_buffer)."""
        for b in self.bufs:
            b.first, b.last = 10 ** 9, -1
        for op in self.ops:
            writes = [op.carry_out_buf] + [o[0] for o in op.outs]
            reads = [op.in_buf, op.carry_in_buf]
            for i in writes:
                if i >= 0:
                    self.bufs[i].first = min(self.bufs[i].first, op.index)
                    self.bufs[i].last = max(self.bufs[i].last, op.index)
            for i in reads:
                if i >= 0:
                    self.bufs[i].last = max(self.bufs[i].last, op.index)
                    self.bufs[i].first = min(self.bufs[i].first, op.index)
        # greedy first-fit over per-image byte ranges; live ranges [first, last] inclusive
        placed: List[Buf] = []
        top = 0
        for b in sorted(self.bufs, key=lambda t: (t.first, -t.bytes_per_image)):
            size = (b.bytes_per_image + 255) // 256 * 256
            busy = sorted((p.offset, p.offset + (p.bytes_per_image + 255) // 256 * 256)
                          for p in placed if keep_buffers or not (p.last < b.first or p.first > b.last))
            off = 0
            for lo, hi in busy:
                if off + size <= lo:
                    break
                off = max(off, hi)
            b.offset = off
            placed.append(b)
            top = max(top, off + size)
        self.workspace_per_image = top
        return self

    # -- ctypes descriptor --------------------------------------------------------------
    def to_desc(self):
        """Returns (f8_model_desc, keepalive) -- keepalive holds every array the descriptor
        points into and must outlive f8_plan_create."""
        n = len(self.ops)
        ops = (C.f8_op * n)()
        keep = [ops]
        for i, op in enumerate(self.ops):
            o = ops[i]
            o.kind = op.kind
            o.cin, o.cout, o.cin_pad, o.cout_pad = op.cin, op.cout, op.cin_pad, op.cout_pad
            o.kh = o.kw = op.k
            o.stride, o.pad = op.stride, op.pad
            o.hin, o.win, o.hout, o.wout = op.hin, op.win, op.hout, op.wout
            o.in_signed = op.in_signed
            o.in_buf = op.in_buf
            if op.weight is not None:
                w = np.ascontiguousarray(op.weight, dtype=np.int32)
                b = np.ascontiguousarray(op.bias, dtype=np.int32)
                keep += [w, b]
                o.weight = w.ctypes.data
                o.bias = b.ctypes.data
            o.carry_in_buf, o.carry_shift = op.carry_in_buf, op.carry_shift
            o.relu = op.relu
            o.carry_out_buf = op.carry_out_buf
            for j in range(2):
                if j < len(op.outs):
                    o.out_buf[j], o.out_shift[j], o.out_signed[j] = op.outs[j]
                else:
                    o.out_buf[j], o.out_shift[j], o.out_signed[j] = -1, 0, 0
            o.out_f32 = op.out_f32
            o.flags = op.flags
        bufs = (C.f8_buffer * max(1, len(self.bufs)))()
        keep.append(bufs)
        for i, b in enumerate(self.bufs):
            bufs[i].bytes_per_image = b.bytes_per_image
            bufs[i].offset_per_image = b.offset
        d = C.f8_model_desc()
        d.abi_version = C.F8_ABI_VERSION
        d.n_ops, d.ops = n, ops
        d.n_buffers, d.buffers = len(self.bufs), bufs
        d.workspace_per_image = self.workspace_per_image
        d.image_h = d.image_w = self.net.image_size
        d.num_classes = self.net.num_classes
        d.head_signed = int(self.net.head.sym)
        return d, keep

    # -- accounting (DESIGN.md / bench.py roofline) ---------------------------------------
    def traffic_per_image(self):
        """Bytes each launch moves per image as planned (reads + writes of activations)."""
        rows = []
        for op in self.ops:
            rd = sum(self.bufs[i].bytes_per_image for i in (op.in_buf, op.carry_in_buf) if i >= 0)
            if op.in_buf == -2:
                rd += op.hin * op.win * 3 * 4
            wr = sum(self.bufs[i].bytes_per_image
                     for i in [op.carry_out_buf] + [o[0] for o in op.outs] if i >= 0)
            if op.out_f32:
                wr += op.cout * 4
            rows.append((op.name, op.kind, rd, wr))
        return rows


def _layer(sd, spec: ConvSpec):
    w = np.asarray(sd[spec.prefix + ".weight"])
    b = np.asarray(sd[spec.prefix + ".bias"])
    fw = int(np.asarray(sd[spec.prefix + ".weight_fraclen"]).reshape(-1)[0])
    fi = int(np.asarray(sd[spec.prefix + ".input_fraclen"]).reshape(-1)[0])
    if tuple(w.shape) != tuple(spec.weight_shape()):
        raise ValueError(f"{spec.prefix}.weight has shape {tuple(w.shape)}, the architecture "
                         f"expects {tuple(spec.weight_shape())}")
    if b.shape != (spec.cout,):
        raise ValueError(f"{spec.prefix}.bias has shape {tuple(b.shape)}, expected ({spec.cout},)")
    return w, b, fw, fi


def _out_hw(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def build_plan(net: NetSpec, sd: Dict[str, np.ndarray], fuse_head: bool = False,
               fuse_tail: bool = False, keep_buffers: bool = False) -> Plan:
    """Lower the integer graph to fused launches.  ``sd`` maps the reference state_dict keys
    to integer arrays (numpy or anything np.asarray accepts).  ``fuse_head``: emit the ResNet
    head conv + ReLU + max-pool as one F8_OP_HEAD_POOL launch (tcgen05 backend)."""
    P = Plan(net)
    S = net.image_size

    def conv_op(spec: ConvSpec, producer: Op, fa_in: int, relu: bool, hin: int, requant=True):
        w, b, fw, fi = _layer(sd, spec)
        dense = not spec.depthwise
        if spec.depthwise and not (spec.k == 3 and spec.pad == 1 and spec.stride in (1, 2)
                                   and spec.cin == spec.cout == spec.groups):
            raise ValueError(f"{spec.prefix}: unsupported grouped convolution")
        op = Op(C.F8_OP_CONV_DENSE if dense else C.F8_OP_CONV_DW, spec.prefix,
                cin=spec.cin, cout=spec.cout, cin_pad=producer.cout_pad, cout_pad=cpad(spec.cout),
                k=spec.k, stride=spec.stride, pad=spec.pad, hin=hin, win=hin,
                hout=_out_hw(hin, spec.k, spec.stride, spec.pad),
                wout=_out_hw(hin, spec.k, spec.stride, spec.pad),
                in_signed=int(spec.sym), weight=w, bias=b, relu=int(relu), fa=fw + fi)
        if requant:
            # int_op_only_fix_quant(x, 8, fi, fa_in, sym) in the producer's epilogue
            op.in_buf = P.image8(producer, fa_in - fi, spec.sym)
        return op

    # ---- input + head (the head conv does not requantise: fix_resnet.py:355-358) ----
    x8 = P.new_buf(S * S * 4, "x:nhwc4")
    conv_in = P.emit(Op(C.F8_OP_CONVERT_INPUT, "input", cin=3, cout=3, cin_pad=4, cout_pad=4,
                        hin=S, win=S, hout=S, wout=S, in_signed=int(net.head.sym), in_buf=-2))
    conv_in.outs.append((x8, 0, int(net.head.sym)))
    head = conv_op(net.head, conv_in, 0, True, S, requant=False)
    head.in_buf = x8
    fused = (fuse_head and net.maxpool and S == 224 and net.head.k == 7 and net.head.stride == 2
             and net.head.pad == 3 and net.head.cin == 3 and net.head.cout == 64)
    if fused:
        # x = self.head[:-1](x); x = self.head[-1](x.float()).int()  (fix_resnet.py:355-362)
        head.kind = C.F8_OP_HEAD_POOL
        head.name = "head.0+maxpool"
        head.hout = head.wout = _out_hw(head.hout, 3, 2, 1)
        head.flags = C.F8_OPF_INT_MAXPOOL if net.maxpool_int else 0      # FXQMaxPool2d, fix_resnet.py:355-356
        P.emit(head)
        cur, fa, hw = head, head.fa, head.hout
    else:
        P.emit(head)
        cur, fa, hw = head, head.fa, head.hout
        if net.maxpool:
            # x = self.head[-1](x.float()).int(), fix_resnet.py:358-359
            mp = Op(C.F8_OP_MAXPOOL, "head.maxpool", cin=cur.cout, cout=cur.cout,
                    cin_pad=cur.cout_pad, cout_pad=cur.cout_pad, k=3, stride=2, pad=1, hin=hw, win=hw,
                    hout=_out_hw(hw, 3, 2, 1), wout=_out_hw(hw, 3, 2, 1), in_buf=P.carry(cur), fa=fa,
                    flags=C.F8_OPF_INT_MAXPOOL if net.maxpool_int else 0)
            P.emit(mp)
            cur, hw = mp, mp.hout

    # ---- blocks ----
    for blk in net.blocks:
        x_op, fa_x, hw_x = cur, fa, hw
        sc_op = None
        if blk.shortcut is not None:
            sc_op = P.emit(conv_op(blk.shortcut, x_op, fa_x, False, hw_x))
            P.carry(sc_op)
        r_op, fa_r, hw_r = x_op, fa_x, hw_x
        nb = len(blk.body)
        for i, spec in enumerate(blk.body):
            relu = (i < nb - 1) or blk.relu_after_last
            op = conv_op(spec, r_op, fa_r, relu, hw_r)
            if i == nb - 1 and (blk.identity or sc_op is not None):
                # IntBlock residual: fix_resnet.py:40-77 / fix_mobilenet_v2.py:34-48
                if blk.identity:
                    op.carry_in_buf, fa_s = P.carry(x_op), fa_x
                else:
                    op.carry_in_buf, fa_s = sc_op.carry_out_buf, sc_op.fa
                op.carry_shift = op.fa - fa_s
                op.relu = int(blk.post_relu)
                op.fa = max(op.fa, fa_s)
            P.emit(op)
            r_op, fa_r, hw_r = op, op.fa, op.hout
        cur, fa, hw = r_op, fa_r, hw_r

    # ---- MBV2 tail: requant (signed) -> 1x1 -> ReLU, fix_mobilenet_v2.py:217-220 ----
    if net.tail is not None:
        t = P.emit(conv_op(net.tail, cur, fa, True, hw))
        cur, fa, hw = t, t.fa, t.hout

    # ---- FXQAvgPool2d + requant + classifier + .float() ----
    if hw != 7:
        raise ValueError(f"final feature map is {hw}x{hw}; FXQAvgPool2d(7) expects 7x7")
    fa_pool = fa + 6                       # shiftnum = round(log2(49)), fix_quant_ops.py:121-122
    if fa_pool > 32:                       # the reference's own assert, fix_quant_ops.py:129
        raise AssertionError("FXQAvgPool2d: output_fraclen <= 32 violated")
    w_fc, b_fc, fw_fc, fi_fc = _layer(sd, net.fc)
    if fuse_tail:
        # FXQAvgPool2d + int_op_only_fix_quant + classifier + .float() in one launch
        # (fix_quant_ops.py:126-134, fix_resnet.py:367-383): the pooled 8-bit vector stays on the SM
        tail = Op(C.F8_OP_POOL_FC, "avgpool+" + net.fc.prefix, cin=net.fc.cin, cout=net.fc.cout,
                  cin_pad=cur.cout_pad, cout_pad=cpad(net.fc.cout), k=hw, stride=1, pad=0, hin=hw, win=hw,
                  hout=1, wout=1, in_signed=int(net.fc.sym), in_buf=P.carry(cur), weight=w_fc, bias=b_fc,
                  fa=fw_fc + fi_fc, out_f32=1)
        tail.outs.append((-1, fa_pool - fi_fc, int(net.fc.sym)))      # requant of the pooled sum, no buffer
        P.emit(tail)
        return P.finalize(keep_buffers)
    pool = Op(C.F8_OP_POOL_REQUANT, "avgpool", cin=cur.cout, cout=cur.cout, cin_pad=cur.cout_pad,
              cout_pad=cur.cout_pad, k=hw, stride=1, pad=0, hin=hw, win=hw, hout=1, wout=1,
              in_buf=P.carry(cur), fa=fa_pool)
    P.emit(pool)
    fc = conv_op(net.fc, pool, fa_pool, False, 1)
    fc.out_f32 = 1
    P.emit(fc)
    return P.finalize(keep_buffers)
