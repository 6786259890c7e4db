"""Algorithmic work of the int_op_only path -- the numerators of the
roofline figures bench.py reports.

Definition: every conv / FC layer reads its input once and writes its output once at
1 B/element (requant fused at the producer); int32 moves only where a residual add needs the
unquantised accumulator (identity-block input: 4 B write + 4 B read; shortcut-conv output:
4 B write + 4 B read); weights are read once per launch as int8.  Pooling traffic, the
float logits and the input repack are excluded.  Element counts use LOGICAL channels.
"""
from . import _capi as C
from .arch import NetSpec


def _hw(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def network_work(net: NetSpec):
    """(int8 ops per image = 2*MACs, algorithmic HBM bytes per image, int8 weight bytes)."""
    macs = act = carry = wbytes = 0
    h = net.image_size

    def layer(spec, hin):
        nonlocal macs, act, wbytes
        ho = _hw(hin, spec.k, spec.stride, spec.pad) if spec.kind == "conv" else 1
        kk = spec.k * spec.k if spec.kind == "conv" else 1
        macs += ho * ho * spec.cout * (spec.cin // spec.groups) * kk
        act += hin * hin * spec.cin + ho * ho * spec.cout
        wbytes += spec.cout * (spec.cin // spec.groups) * kk
        return ho

    h = layer(net.head, h)
    if net.maxpool:
        h = _hw(h, 3, 2, 1)
    for blk in net.blocks:
        hin = h
        if blk.shortcut is not None:
            layer(blk.shortcut, hin)
        for spec in blk.body:
            h = layer(spec, h)
        if blk.identity:
            carry += 8 * hin * hin * blk.body[0].cin
        elif blk.shortcut is not None:
            carry += 8 * h * h * blk.body[-1].cout
    if net.tail is not None:
        h = layer(net.tail, h)
    macs += net.fc.cin * net.fc.cout
    act += net.fc.cin + net.fc.cout
    wbytes += net.fc.cin * net.fc.cout
    return 2 * macs, act + carry, wbytes


def op_work(plan):
    """Per launch of ``plan``: dict(name, kind, ops, bytes_per_image, weight_bytes) with the
    same definition, attributed to the launch that moves the bytes."""
    read_as_carry = {op.carry_in_buf for op in plan.ops if op.carry_in_buf >= 0}
    rows = []
    for op in plan.ops:
        el_in = op.hin * op.win * op.cin
        el_out = op.hout * op.wout * op.cout
        if op.kind == C.F8_OP_HEAD_POOL:
            # conv at the un-pooled resolution; only the pooled tensor is written
            hc = _hw(op.hin, op.k, op.stride, op.pad)
            macs = hc * hc * op.cout * op.cin * op.k * op.k
            wb = op.cout * op.cin * op.k * op.k
        elif op.kind == C.F8_OP_CONV_DENSE:
            macs = el_out * op.cin * op.k * op.k
            wb = op.cout * op.cin * op.k * op.k
        elif op.kind == C.F8_OP_CONV_DW:
            macs = el_out * 9
            wb = op.cout * 9
        elif op.kind == C.F8_OP_POOL_FC:
            macs = op.cout * op.cin
            wb = op.cout * op.cin
        else:
            macs = wb = 0
        b = 0
        if op.kind == C.F8_OP_HEAD_POOL:
            hc = _hw(op.hin, op.k, op.stride, op.pad)
            b = el_in + hc * hc * op.cout       # SURVEY definition: conv in + conv out, pool excluded
        elif op.kind == C.F8_OP_POOL_FC:
            b = op.cin + op.cout              # the classifier's 8-bit input + its outputs (pooling excluded)
        elif op.kind in (C.F8_OP_CONV_DENSE, C.F8_OP_CONV_DW):
            b = el_in + el_out
            if op.carry_in_buf >= 0:
                b += 4 * el_out
        if op.carry_out_buf in read_as_carry:
            b += 4 * el_out
        rows.append(dict(name=op.name, kind=op.kind, ops=2 * macs, bytes_per_image=b,
                         weight_bytes=wb))
    return rows
