"""Runs the UNMODIFIED reference (/root/reference) IntModel on CPU -- authoring container only.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box; nothing under
tests -m gpu, smoke() or bench.py imports this module.  It exists to (a) validate the C
restatement in oracle/ against the real reference and (b) generate the golden vectors
committed under tests/golden/ (driver: tests/golden/make_golden.py).

The reference keeps its config in a module-level singleton read from sys.argv at import
(/root/reference/myutils/config.py:152-178), so one process can host one architecture:
use it as a script,

    python -m oracle.ref_harness <arch> <in.npz> <out.npz> [flag=0|1 ...]

in.npz : 'x' int32 [N,3,224,224] + every state_dict tensor under its reference key.
out.npz: 'logits' float32 [N,1000], 'keys' (state_dict key order of the reference IntModel),
         and one '<prefix>:in8' / '<prefix>:acc' int32 array per int layer (forward hooks).

Two source-untouched shims: the sys.argv config shim and the
requires_grad=False parameter shim needed on torch >= 2 by fix_quant_ops.py:705-709.
"""
import contextlib
import importlib
import os
import sys

REF_ROOT = "/root/reference"
CONFIGS = {
    "resnet18": "apps/imagenet/resnet18/conventional/res18_fix_quant_test_int_op_only.yml",
    "resnet50": "apps/imagenet/resnet50/tiny_finetuning/"
                "res50_fix_quant_ptcv_pretrained_test_int_op_only_on_cpu.yml",
    "mobilenet_v1": "apps/imagenet/mobilenetv1/conventional/"
                    "mbv1_fix_quant_test_int_op_only_on_cpu.yml",
    "mobilenet_v2": "apps/imagenet/mobilenetv2/conventional/"
                    "mbv2_fix_quant_test_int_op_only_on_cpu.yml",
}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def build_int_model(arch, float_state_dict=None, keep_grid_search=False, flag_overrides=None):
    """Return (IntModel, FLAGS) built the way fix_train.py:258-296 + :930-934 does, on CPU.
    ``float_state_dict``: loaded into the float-sim Model before ``int_model()`` converts it (the
    export golden vectors, tests/golden/make_export_golden.py)."""
    import torch
    import torch.nn as nn

    sys.dont_write_bytecode = True
    sys.argv = ["x", "app:" + os.path.join(REF_ROOT, CONFIGS[arch]), "bs:1"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from myutils.config import FLAGS
    from models.fix_quant_ops import ReLUClipFXQConvBN, ReLUClipFXQLinear

    @contextlib.contextmanager
    def nograd_param_shim():
        oc, ol = nn.Conv2d, nn.Linear

        class C(oc):
            def __init__(s, *a, **k):
                super().__init__(*a, **k)
                for p in s.parameters():
                    p.requires_grad_(False)

        class L(ol):
            def __init__(s, *a, **k):
                super().__init__(*a, **k)
                for p in s.parameters():
                    p.requires_grad_(False)

        nn.Conv2d, nn.Linear = C, L
        try:
            yield
        finally:
            nn.Conv2d, nn.Linear = oc, ol

    for k, v in (flag_overrides or {}).items():    # config variants of the export golden vectors
        setattr(FLAGS, k, v)
    if getattr(FLAGS, "format_grid_search", False) and not keep_grid_search:
        # grid search only affects the *values* int_model() exports, which we overwrite
        # with the synthetic state dict; skip its cost (fix_quant_ops.py:17-27).
        FLAGS.format_grid_search = False
        FLAGS.format_from_metric = True
    model = importlib.import_module(FLAGS.model).Model(FLAGS.num_classes)
    for m in model.modules():  # mirrors fix_train.py:270-295
        if isinstance(m, (ReLUClipFXQConvBN, ReLUClipFXQLinear)):
            m.set_weight_format(FLAGS.weight_format)
            m.set_input_format(FLAGS.input_format)
            m.rescale_type = getattr(FLAGS, "rescale_type", "constant")
            m.set_alpha()
            m.floating = getattr(FLAGS, "floating_model", False)
            m.floating_wo_clip = getattr(FLAGS, "floating_wo_clip", False)
            m.format_type = getattr(FLAGS, "format_type", None)
            m.format_from_metric = getattr(FLAGS, "format_from_metric", False)
            m.metric = getattr(FLAGS, "metric", None)
            m.format_grid_search = getattr(FLAGS, "format_grid_search", False)
            m.set_metric_func()
            m.register_input_format(FLAGS.input_format,
                                    momentum=getattr(FLAGS, "momentum_for_metric", 0.1))
            m.no_clipping = getattr(FLAGS, "no_clipping", False)
            m.input_fraclen_sharing = getattr(FLAGS, "input_fraclen_sharing", False)
            m.quant_bias = getattr(FLAGS, "quant_bias", False)
            m.int_infer = getattr(FLAGS, "int_infer", False)
        if isinstance(m, ReLUClipFXQConvBN):
            m.rescale_forward = getattr(FLAGS, "rescale_forward_conv", False)
        if isinstance(m, ReLUClipFXQLinear):
            m.rescale_forward = getattr(FLAGS, "rescale_forward", True)
    model.eval()
    if float_state_dict is not None:       # a float-sim checkpoint (fix_train.py:877-891)
        missing, unexpected = model.load_state_dict(float_state_dict, strict=False)
        assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    model.apply(lambda m: setattr(m, "int_op_only", True))       # fix_train.py:932
    with nograd_param_shim():
        im = model.int_model().cpu()                             # fix_train.py:933
    im.apply(lambda m: setattr(m, "int_op_only", True))          # fix_train.py:934
    im.eval()
    return im, FLAGS


def run(arch, x_np, sd_np, capture=True, flag_overrides=None):
    """Load the synthetic state dict into the reference IntModel and run its forward."""
    import numpy as np
    import torch
    import torch.nn as nn

    im, FLAGS = build_int_model(arch, flag_overrides=flag_overrides)
    keys = list(im.state_dict().keys())
    ref_sd = im.state_dict()
    new_sd = {}
    for k in keys:
        t = torch.from_numpy(np.ascontiguousarray(sd_np[k])).to(torch.int32)
        new_sd[k] = t.reshape(ref_sd[k].shape)
    im.load_state_dict(new_sd)
    out = {}
    if capture:
        for name, m in im.named_modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                def pre(mod, inp, name=name):
                    out[name + ":in8"] = inp[0].detach().numpy().astype(np.int32).copy()

                def post(mod, inp, res, name=name):
                    out[name + ":acc"] = res.detach().numpy().astype(np.int32).copy()
                m.register_forward_pre_hook(pre)
                m.register_forward_hook(post)
    x = torch.from_numpy(np.ascontiguousarray(x_np)).to(torch.int32)
    head_fi = int(im.head[0].input_fraclen.item())
    x.output_fraclen = head_fi if getattr(FLAGS, "normalize", False) else 8
    with torch.no_grad():
        logits = im(x)
    sym = {name: bool(getattr(m, "input_symmetric", False)) for name, m in im.named_modules()
           if isinstance(m, (nn.Conv2d, nn.Linear))}
    return logits.numpy(), keys, out, sym, im


def main(argv):
    import numpy as np
    arch, inp, outp = argv[1], argv[2], argv[3]
    # optional FLAGS overrides, e.g. quant_maxpool=1 (FXQMaxPool2d head pool)
    overrides = {kv.split("=")[0]: bool(int(kv.split("=")[1])) for kv in argv[4:]}
    data = np.load(inp)
    sd = {k: data[k] for k in data.files if k != "x"}
    logits, keys, cap, sym, _ = run(arch, data["x"], sd, flag_overrides=overrides)
    cap = {k: v for k, v in cap.items()}
    np.savez(outp, logits=logits, keys=np.array(keys),
             sym_names=np.array(list(sym.keys())),
             sym_vals=np.array(list(sym.values()), dtype=np.int32), **cap)


if __name__ == "__main__":
    main(sys.argv)
