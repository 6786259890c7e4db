"""Serialised engine file ("offline export -> packed int8 engine file").

The reference persists the integer model as a pickled ``{'model': <bound method state_dict>}``
(fix_train.py:940-947) holding int32 tensors -- 4 bytes per 8-bit weight -- and needs its Python
module tree to rebuild the graph.  The engine file is self-contained instead: one ``.npz`` with

    meta                 JSON: format version, architecture name, head signedness (FLAGS.normalize),
                         quant_maxpool, classes, image size, layer order
    <prefix>.weight      int8, reference layout [O, C/g, kh, kw] | [O, K]   (values are in [-127, 127])
    <prefix>.bias        int32 [O]
    <prefix>.weight_fraclen / .input_fraclen     int32, the reference's shapes

``load_engine(path)`` returns a ready Engine (the per-layer tensor-core tile packing happens in
f8_plan_create, so the file does not depend on the kernels' internal layouts); ``load_state_dict(path)``
returns the reference-layout int32 state_dict again, bit for bit."""
import json

import numpy as np

from .arch import graph_for

FORMAT = "f8net-b200-engine"
VERSION = 1


def save_engine(path, state_dict, arch=None, head_signed=False, quant_maxpool=False):
    """Write ``state_dict`` (reference IntModel layout, or the bound method the reference pickles) as
    an engine file.  Raises if a weight does not fit 8 bits (not an F8Net integer model)."""
    from .engine import _to_numpy_sd, infer_arch
    sd = state_dict() if callable(state_dict) else state_dict
    sd = _to_numpy_sd(sd)
    arch = arch or infer_arch(sd)
    net = graph_for(arch, bool(head_signed), quant_maxpool=bool(quant_maxpool))
    out, order = {}, []
    for L in net.convs():
        p = L.prefix
        w = np.asarray(sd[p + ".weight"])
        if tuple(w.shape) != tuple(L.weight_shape()):
            raise ValueError(f"{p}.weight has shape {tuple(w.shape)}, {arch} expects {tuple(L.weight_shape())}")
        if w.min() < -127 or w.max() > 127:
            raise ValueError(f"{p}.weight leaves the 8-bit range [{w.min()}, {w.max()}]")
        out[p + ".weight"] = w.astype(np.int8)
        out[p + ".bias"] = np.asarray(sd[p + ".bias"], dtype=np.int32)
        out[p + ".weight_fraclen"] = np.asarray(sd[p + ".weight_fraclen"], dtype=np.int32)
        out[p + ".input_fraclen"] = np.asarray(sd[p + ".input_fraclen"], dtype=np.int32)
        order.append(p)
    meta = {"format": FORMAT, "version": VERSION, "arch": arch, "head_signed": bool(head_signed),
            "quant_maxpool": bool(quant_maxpool), "num_classes": net.num_classes,
            "image_size": net.image_size, "layers": order}
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    with open(path, "wb") as f:
        np.savez(f, **out)
    return meta


def _read(path):
    z = np.load(path)
    if "meta" not in z.files:
        raise ValueError(f"{path}: not an f8net-b200 engine file")
    meta = json.loads(bytes(z["meta"]).decode())
    if meta.get("format") != FORMAT or meta.get("version") != VERSION:
        raise ValueError(f"{path}: unsupported engine file {meta.get('format')} v{meta.get('version')}")
    return z, meta


def load_state_dict(path):
    """(meta, reference-layout int32 state_dict in the reference's key order)."""
    z, meta = _read(path)
    sd = {}
    for p in meta["layers"]:
        sd[p + ".weight"] = z[p + ".weight"].astype(np.int32)
        sd[p + ".bias"] = z[p + ".bias"]
        sd[p + ".weight_fraclen"] = z[p + ".weight_fraclen"]
        sd[p + ".input_fraclen"] = z[p + ".input_fraclen"]
    return meta, sd


def load_engine(path, device=None, chunk=256, backend=None):
    """Engine from an engine file (needs a CUDA device, like ``compile``)."""
    from .engine import compile as _compile
    meta, sd = load_state_dict(path)
    return _compile(sd, arch=meta["arch"], head_signed=meta["head_signed"], device=device, chunk=chunk,
                    backend=backend, quant_maxpool=meta["quant_maxpool"])
