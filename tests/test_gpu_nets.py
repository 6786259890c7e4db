"""Whole-network parity on the GPU through the reference-facing call surface
(f8net_b200.compile(...)(x) == IntModel.forward(x)):

 * against the committed golden logits produced by the UNMODIFIED reference
   (tests/golden/*.npz, N=2, calibrated + adversarial fixtures);
 * against the CPU oracle on fresh seeded inputs (ragged batches / chunks);
 * at BASELINE.json's full batch (256 per GPU) through size-independent properties:
   chunking invariance, batch-permutation equivariance, determinism, and agreement of the
   engine-native NHWC u8 input with the reference's int32 NCHW input.

Integer path => bit-exact equality everywhere (logits are float32 holding exact int32s)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import f8net_b200  # noqa: E402
from f8net_b200 import synth  # noqa: E402
from oracle import nets  # noqa: E402

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ARCHS = list(synth.HEAD_SIGNED)


def _engine(arch, sd, **kw):
    return f8net_b200.compile(sd, arch=arch, head_signed=synth.HEAD_SIGNED[arch], **kw)


def _nhwc4(x, signed):
    n, _, h, w = x.shape
    out = np.zeros((n, h, w, 4), dtype=np.int8 if signed else np.uint8)
    out[..., :3] = x.transpose(0, 2, 3, 1)
    return out


@pytest.mark.parametrize("backend", [pytest.param(0, id="imma"), pytest.param(1, id="tcgen05")])
@pytest.mark.parametrize("family", ["calibrated", "edge"])
@pytest.mark.parametrize("arch", ARCHS)
def test_golden_logits_from_the_reference(cuda, f8lib, arch, family, backend):
    hs = synth.HEAD_SIGNED[arch]
    if family == "calibrated":
        sd, x = synth.make_state_dict(arch, hs), synth.make_input(arch, 2, hs)
        gold = np.load(os.path.join(GOLD, f"{arch}_n2.npz"))
    else:
        sd, x = synth.make_edge_state_dict(arch, hs), synth.make_input(arch, 2, hs, seed=777)
        gold = np.load(os.path.join(GOLD, f"edge_{arch}_n2.npz"))
    eng = _engine(arch, sd, backend=backend)
    # the reference's own call: CPU int32 NCHW tensor in, float32 logits out
    xt = torch.from_numpy(x)
    xt.output_fraclen = 8
    y = eng(xt)
    assert y.dtype == torch.float32 and tuple(y.shape) == (2, 1000) and not y.is_cuda
    assert np.array_equal(y.numpy().astype(np.int64), gold["logits"].astype(np.int64))
    # same through device tensors, and through the engine-native NHWC 8-bit input
    y2 = eng(xt.cuda())
    assert torch.equal(y2.cpu(), y)
    y3 = eng.run_device(torch.from_numpy(_nhwc4(x, hs)).cuda())
    assert torch.equal(y3.cpu(), y)


@pytest.mark.parametrize("name", synth.TRAINED)
def test_trained_fraclen_family_golden(cuda, f8lib, name):
    """The per-layer formats of the networks the reference's authors trained (fraclen_visual/*.out;
    fi 1..8, fw 0..7): golden logits from the unmodified reference (make_variant_golden.py)."""
    from util import trained_fixture
    arch, hs, sd, x, gold = trained_fixture(name)
    for backend in (0, 1):
        eng = f8net_b200.compile(sd, arch=arch, head_signed=hs, backend=backend)
        y = eng(torch.from_numpy(x))
        assert np.array_equal(y.numpy().astype(np.int64), gold["logits"].astype(np.int64)), (name, backend)


@pytest.mark.parametrize("backend", [pytest.param(0, id="imma"), pytest.param(1, id="tcgen05")])
@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_both_head_pools_golden(cuda, f8lib, arch, backend):
    """FLAGS.quant_maxpool (fix_resnet.py:331-334, :355-359): FXQMaxPool2d's integer max against
    nn.MaxPool2d on floats, on a fixture where 99 % of the reference's logits differ between them."""
    from util import qmaxpool_fixture
    hs, sd, x, gold = qmaxpool_fixture(arch)
    xt = torch.from_numpy(x)
    yf = f8net_b200.compile(sd, arch=arch, head_signed=hs, backend=backend)(xt)
    yi = f8net_b200.compile(sd, arch=arch, head_signed=hs, backend=backend, quant_maxpool=True)(xt)
    assert np.array_equal(yf.numpy().astype(np.int64), gold["logits_float_pool"].astype(np.int64))
    assert np.array_equal(yi.numpy().astype(np.int64), gold["logits"].astype(np.int64))


@pytest.mark.parametrize("arch", ARCHS)
def test_oracle_parity_ragged_batch_and_chunks(cuda, f8lib, arch):
    hs = synth.HEAD_SIGNED[arch]
    sd = synth.make_state_dict(arch, hs)
    x = synth.make_input(arch, 7, hs, seed=4242)
    want = nets.forward(arch, sd, x, hs)
    xt = torch.from_numpy(x).cuda()
    for chunk in (7, 3, 1):
        eng = _engine(arch, sd, chunk=chunk)
        y = eng(xt).cpu().numpy()
        assert np.array_equal(y, want), (arch, chunk)
        assert eng.launches(7) == len(eng.plan.ops) * -(-7 // chunk)


@pytest.mark.parametrize("arch", ARCHS)
def test_full_batch_properties(cuda, f8lib, arch):
    """BASELINE.json batch (256 images on one GPU): properties that do not need the oracle
    at full size, plus an oracle spot check of 4 images scattered through the batch."""
    hs = synth.HEAD_SIGNED[arch]
    n = 256
    sd = synth.make_state_dict(arch, hs)
    x = synth.make_input(arch, n, hs)
    xt = torch.from_numpy(x).cuda()
    eng = _engine(arch, sd, chunk=32)
    y = eng(xt)
    torch.cuda.synchronize()
    # determinism
    assert torch.equal(eng(xt), y)
    # chunking invariance (ragged last chunk)
    assert torch.equal(_engine(arch, sd, chunk=n)(xt), y)
    assert torch.equal(_engine(arch, sd, chunk=48)(xt), y)
    # images are independent: permuting the batch permutes the logits
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).cuda()
    assert torch.equal(eng(xt[perm].contiguous()), y[perm])
    # engine-native layout agrees with the reference layout
    assert torch.equal(eng.run_device(torch.from_numpy(_nhwc4(x, hs)).cuda()), y)
    # logits are exact integers and not degenerate (every image gets its own logits)
    yc = y.cpu().numpy()
    assert np.array_equal(yc, np.rint(yc)) and np.unique(yc, axis=0).shape[0] == n
    idx = [0, 97, 200, 255]
    want = nets.forward(arch, sd, x[idx], hs)
    assert np.array_equal(yc[idx], want)


@pytest.mark.parametrize("arch", ARCHS)
def test_full_batch_every_image_against_the_oracle(cuda, f8lib, arch):
    """The bench configuration itself -- 256 images, one pass (chunk 256) -- with EVERY image's logits compared with
    the CPU oracle (the full batch was covered by a 4-image spot check plus invariances).  The oracle
    runs 32 images at a time to bound host memory; a few seconds per network on the GPU box's cores."""
    hs = synth.HEAD_SIGNED[arch]
    n = 256
    sd = synth.make_state_dict(arch, hs)
    x = synth.make_input(arch, n, hs, seed=31337)
    y = _engine(arch, sd, chunk=n)(torch.from_numpy(x).cuda()).cpu().numpy()
    for i0 in range(0, n, 32):
        want = nets.forward(arch, sd, x[i0:i0 + 32], hs)
        assert np.array_equal(y[i0:i0 + 32], want), (arch, i0)


def test_compile_from_module_like_object_and_bound_method(cuda, f8lib):
    """compile() accepts the state_dict, or the bound method the reference pickles
    (fix_train.py:946 saves {'model': model_wrapper.state_dict} without calling it)."""
    sd = synth.make_state_dict("mobilenet_v1")
    tsd = synth.to_torch_state_dict(sd)

    class Holder:
        def state_dict(self):
            return tsd

    x = torch.from_numpy(synth.make_input("mobilenet_v1", 2))
    a = f8net_b200.compile(tsd)(x)                       # arch inferred from the keys
    b = f8net_b200.compile(Holder().state_dict)(x)       # bound method
    assert torch.equal(a, b)
    want = nets.forward("mobilenet_v1", sd, x.numpy())
    assert np.array_equal(a.numpy(), want)


def test_input_validation(cuda, f8lib):
    eng = _engine("mobilenet_v1", synth.make_state_dict("mobilenet_v1"))
    with pytest.raises(TypeError):
        eng(torch.zeros((1, 3, 224, 224), dtype=torch.float16).cuda())
    with pytest.raises(TypeError):
        eng(torch.zeros((1, 3, 112, 112), dtype=torch.float32).cuda())              # wrong image size
    with pytest.raises(TypeError):
        eng.run_device(torch.zeros((1, 224, 224, 4), dtype=torch.int8).cuda())   # head is unsigned
    bad = torch.full((1, 3, 224, 224), 300, dtype=torch.int32)
    with pytest.raises(ValueError, match="outside"):
        eng(bad, strict=True)
    # the float tensor of forward_loss: the reference asserts input >= 0 (fix_train.py:689)
    with pytest.raises(ValueError, match="outside"):
        eng(torch.full((1, 3, 224, 224), -0.25), strict=True)
    with pytest.raises(ValueError, match="outside"):
        eng(torch.full((1, 3, 224, 224), 1.5).cuda(), strict=True)
    eng(torch.rand((1, 3, 224, 224)), strict=True)


@pytest.mark.parametrize("switch,value,arch", [
    ("F8_MC", "0", "resnet18"), ("F8_MC", "4", "resnet18"), ("F8_MC_GENERIC", "0", "resnet50"),
    ("F8_CPA", "1", "mobilenet_v2"), ("F8_PDL", "0", "resnet18"), ("F8_GATHER_NO_TMA", "1", "mobilenet_v2"),
    ("F8_NO_CONV1X1_RES", "1", "mobilenet_v2"), ("F8_DW_CUDA_CORE", "1", "mobilenet_v1")])
def test_switchable_kernel_paths_stay_exact(cuda, f8lib, switch, value, arch):
    """Kernel variants kept behind environment switches: the 3x3 kernel as single CTAs / clusters of four instead of
    pairs sharing the weight stream (F8_MC=0|4; F8_MC_GENERIC for the residual launches alone), the cp.async operand
    loader of the point-wise kernel (F8_CPA=1), plain stream order instead of programmatic dependent launch
    (F8_PDL=0), the gather kernel instead of the TMA / resident-weight point-wise kernels (F8_GATHER_NO_TMA,
    F8_NO_CONV1X1_RES), the CUDA-core depthwise kernel instead of the tensor-core one (F8_DW_CUDA_CORE).  None of them
    may change a bit.  The library reads the switches once, so each runs in its own process: odd batch (ragged
    halves / tiles) against the oracle."""
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, f8net_b200\n"
        "from f8net_b200 import synth\n"
        "from oracle import nets\n"
        f"arch = {arch!r}\n"
        "hs = synth.HEAD_SIGNED[arch]\n"
        "sd = synth.make_state_dict(arch, hs)\n"
        "x = synth.make_input(arch, 5, hs, seed=11)\n"
        "eng = f8net_b200.compile(sd, arch=arch, head_signed=hs, chunk=5)\n"
        "y = eng(torch.from_numpy(x).cuda()).cpu().numpy()\n"
        "assert np.array_equal(y, nets.forward(arch, sd, x, hs)), 'logits differ'\n"
        "print('exact')\n")
    env = dict(os.environ)
    env[switch] = value
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "exact" in r.stdout, r.stderr[-2000:]
