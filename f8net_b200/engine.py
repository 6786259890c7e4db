"""Reference-facing host API: a drop-in for ``IntModel.forward`` on the int_op_only path.

    engine = f8net_b200.compile(int_model)                    # a reference IntModel, or
    engine = f8net_b200.compile(state_dict, arch="resnet18")  # its state_dict (+ arch name)
    logits = engine(x)      # x: int32 [N,3,224,224] NCHW, values in the head's 8-bit range
    logits = engine(xf)     # xf: the DataLoader's float32 [N,3,224,224] -- forward_loss's
                            # integerisation (fix_train.py:676-692) then runs on the device
    logits = engine(img)    # img: decoded uint8 [N,224,224,3] -- ToTensor + Normalize + the
                            # integerisation folded into one table lookup on the device

mirrors ``output = model(input)`` (/root/reference/fix_train.py:693) for the IntModel built
at /root/reference/fix_train.py:930-935: same input tensor, same float32 ``[N, 1000]`` output
holding the exact int32 logits (/root/reference/models/fix_resnet.py:383).  All arithmetic
runs in libf8b200.so on the GPU; torch is used for device memory and streams only.
"""
import ctypes
from typing import Optional

import numpy as np

from . import _capi as C
from .arch import ARCHS, NetSpec, graph_for, graph_from_module
from .planner import Plan, build_plan


def _to_numpy_sd(sd):
    out = {}
    for k, v in sd.items():
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def infer_arch(sd):
    """Architecture name from the state_dict key set."""
    keys = set(sd.keys())
    if "tail.0.weight" in keys:
        return "mobilenet_v2"
    if not any(".shortcut." in k for k in keys) and "stage_4_layer_0.body.0.weight" in keys:
        return "mobilenet_v1"
    n_blocks = len({k.split(".")[0] for k in keys if k.startswith("stage_")})
    bottleneck = any(k.endswith(".body.4.weight") for k in keys)
    table = {(8, False): 18, (16, False): 34, (16, True): 50, (33, True): 101, (50, True): 152}
    depth = table.get((n_blocks, bottleneck))
    if depth is None:
        raise ValueError("cannot infer the architecture from the state_dict; pass arch=")
    return f"resnet{depth}"


class InputRangeError(ValueError):
    """An input tensor held values outside the head's 8-bit range (see ``Engine.check_input_range``)."""


class Engine:
    """One compiled plan on one GPU.  Stateless at inference like the reference's IntModel:
    the only mutable state is scratch memory, so use one Engine per stream."""

    def __init__(self, net: NetSpec, state_dict, device=None, chunk: int = 256, backend=None,
                 keep_buffers: bool = False, fuse_tail=None, range_check: bool = True):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("f8net_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self._torch = torch
        self.lib = C.lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None
                                   else torch.device(device).index or 0)
        self.net = net
        if backend is None:
            backend = 1 if self.lib.f8_has_umma(self.device.index) else 0
        # the fused head conv + max-pool launch exists on the tcgen05 backend only
        # (the fused tail launch -- pool + requant + classifier -- likewise)
        if fuse_tail is None:
            fuse_tail = int(backend) == 1
        self.plan: Plan = build_plan(net, _to_numpy_sd(state_dict), fuse_head=(int(backend) == 1),
                                     fuse_tail=bool(fuse_tail), keep_buffers=keep_buffers)
        self.keep_buffers = bool(keep_buffers)
        self.range_check = bool(range_check)
        self._last = None                      # (n, chunk) of the most recent run_device
        self.chunk = int(chunk)
        desc, keep = self.plan.to_desc()
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            C.check(self.lib.f8_plan_create(ctypes.byref(desc), self.device.index,
                                            ctypes.byref(handle)))
        del keep
        self._h = handle
        self.set_backend(backend)
        self._ws = None
        self._stage = None
        self._logits = None
        self.head_fraclen = int(np.asarray(state_dict["head.0.input_fraclen"]).reshape(-1)[0]) \
            if "head.0.input_fraclen" in state_dict else None
        # forward_loss's input preparation: FLAGS.normalize <=> signed head (fix_resnet.py:437-438)
        self.set_input_prep(bool(net.head.sym))

    IMAGENET_MEAN = (0.485, 0.456, 0.406)      # fix_train.py:303-304
    IMAGENET_STD = (0.229, 0.224, 0.225)

    def set_input_prep(self, normalize: bool, mean=None, std=None, fraclen=None):
        """How float32 / uint8 inputs become the head's 8-bit integers (fix_train.py:676-692):
        normalize False: (255 x).round(); True: clamp(round(x 2^fl), -127, 127) with
        fl = head.input_fraclen, on x = (p / 255 - mean) / std for uint8 pixels p."""
        if normalize:
            mean = self.IMAGENET_MEAN if mean is None else mean
            std = self.IMAGENET_STD if std is None else std
            if fraclen is None:
                fraclen = self.head_fraclen
            if fraclen is None:
                raise ValueError("normalize=True needs head.0.input_fraclen (or fraclen=)")
        else:
            mean, std, fraclen = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), 8
        m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
        s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
        with self._torch.cuda.device(self.device):
            C.check(self.lib.f8_plan_set_input_prep(self._h, int(bool(normalize)), int(fraclen), m3, s3))
        self.input_prep = {"normalize": bool(normalize), "mean": tuple(mean), "std": tuple(std),
                           "fraclen": int(fraclen)}

    # ------------------------------------------------------------------------------------
    def set_backend(self, backend: int):
        """0 = mma.sync IMMA kernels, 1 = tcgen05 (UMMA + TMA) where the shape allows."""
        C.check(self.lib.f8_plan_set_backend(self._h, int(backend)))
        self.backend = int(backend)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.f8_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_per_image(self):
        return self.plan.workspace_per_image

    def launches(self, n, layout=C.F8_IN_NCHW_I32, chunk=None):
        return int(self.lib.f8_plan_launch_count(self._h, layout, int(n), int(chunk or self.chunk)))

    def _grow(self, attr, need, dtype):
        """Scratch tensors grow by replacement.  The caching allocator only knows the allocation
        stream, while the library enqueues on whatever stream the caller passes, so work still in
        flight could see the freed block handed to someone else: drain the device before the old
        tensor is dropped (growth is rare: first call, or a larger batch / chunk)."""
        cur = getattr(self, attr)
        if cur is None or cur.numel() < need:
            if cur is not None:
                self._torch.cuda.synchronize(self.device)
            cur = self._torch.empty(need, dtype=dtype, device=self.device)
            setattr(self, attr, cur)
        return cur

    def _workspace(self, chunk):
        return self._grow("_ws", self.plan.workspace_per_image * chunk, self._torch.uint8)

    def _layout_of(self, x):
        torch = self._torch
        S = self.net.image_size
        shape = tuple(x.shape[1:])
        if x.dtype == torch.int32 and shape == (3, S, S):
            return C.F8_IN_NCHW_I32
        if x.dtype == torch.float32 and shape == (3, S, S):
            return C.F8_IN_NCHW_F32
        if x.dtype == torch.uint8 and shape == (S, S, 3):
            return C.F8_IN_NHWC3_U8
        if x.dtype in (torch.uint8, torch.int8) and shape == (S, S, 4):
            want = torch.int8 if self.net.head.sym else torch.uint8
            if x.dtype != want:
                raise TypeError(f"head conv is {'signed' if self.net.head.sym else 'unsigned'}: "
                                f"expected {want}, got {x.dtype}")
            return C.F8_IN_NHWC4_8
        raise TypeError(f"expected int32 / float32 [N,3,{S},{S}], uint8 [N,{S},{S},3] or 8-bit "
                        f"[N,{S},{S},4]; got {x.dtype} {tuple(x.shape)}")

    # ------------------------------------------------------------------------------------
    def run_device(self, x, out=None, chunk=None, stream=None):
        """x: CUDA tensor: int32 [N,3,H,W] (the reference's tensor), float32 [N,3,H,W] (the
        DataLoader's tensor, integerised on the device), uint8 [N,H,W,3] (decoded pixels) or
        8-bit [N,H,W,4] (engine-native NHWC, channel 3 zero).  Enqueues on the current stream
        (or ``stream``) and returns the float32 [N, classes] logits tensor, no sync."""
        torch = self._torch
        if x.device != self.device:
            raise ValueError(f"input is on {x.device}, the engine on {self.device}")
        layout = self._layout_of(x)
        if not x.is_contiguous():
            x = x.contiguous()
        n = x.shape[0]
        if out is None:
            out = torch.empty((n, self.net.num_classes), dtype=torch.float32, device=self.device)
        elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < n * self.net.num_classes:
            raise ValueError("out must be a contiguous float32 tensor of at least [N, classes]")
        if n == 0:
            return out
        chunk = min(int(chunk or self.chunk), n)
        ws = self._workspace(chunk)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        # asynchronous: an out-of-range input of an EARLIER call surfaces here (or in check_input_range)
        self._raise_if_out_of_range("an earlier asynchronous call")
        C.check(self.lib.f8_plan_run(self._h, x.data_ptr(), layout, n, out.data_ptr(),
                                     ws.data_ptr(), ws.numel(), chunk, st.cuda_stream))
        self._last = (n, chunk)
        return out

    def _raise_if_out_of_range(self, who):
        if self.lib.f8_plan_input_range(self._h, 1) and self.range_check:
            lo, hi = (-128, 127) if self.net.head.sym else (0, 255)
            raise InputRangeError(
                f"{who} had input values outside the head's 8-bit range [{lo}, {hi}] (or NaN): the "
                f"reference's head conv consumes the full int32 (fix_resnet.py:355) and forward_loss asserts "
                f"input >= 0 (fix_train.py:689); the engine computed from the low bytes, its logits differ")

    def check_input_range(self, stream=None):
        """The always-on input range check for asynchronous calls (``run_device`` / ``__call__`` on a CUDA
        tensor / ``run_host(sync=False)``): waits for ``stream`` (default: the current one) and raises
        InputRangeError if any input since the last check lay outside the head's 8 bits.  Synchronous
        ``run_host`` calls raise by themselves.  ``Engine(range_check=False)`` never raises."""
        torch = self._torch
        (stream if stream is not None else torch.cuda.current_stream(self.device)).synchronize()
        self._raise_if_out_of_range("an input since the last check")

    def read_buffer(self, index):
        """Debug / parity aid: plan buffer ``index`` (see ``plan.bufs``) as the most recent single-pass
        ``run_device`` left it, as a uint8 numpy array of n * bytes_per_image bytes.  Needs an Engine
        built with ``keep_buffers=True`` (otherwise dead buffers are reused by later layers)."""
        if not self.keep_buffers:
            raise RuntimeError("read_buffer needs compile(..., keep_buffers=True)")
        if self._last is None or self._last[0] > self._last[1]:
            raise RuntimeError("read_buffer: run one pass first (n <= chunk)")
        n, chunk = self._last
        dst = np.empty(self.plan.bufs[index].bytes_per_image * n, dtype=np.uint8)
        st = self._torch.cuda.current_stream(self.device)
        C.check(self.lib.f8_plan_read_buffer(self._h, int(index), n, chunk, self._ws.data_ptr(),
                                             dst.ctypes.data, dst.nbytes, st.cuda_stream))
        return dst

    def run_host(self, x, chunk=None, out=None, sync=True, stream=None):
        """x: CPU tensor (int32 NCHW or 8-bit NHWC4; pinned for full PCIe speed).  Copies in,
        runs, copies the logits back through f8_plan_run_host and returns a CPU tensor.
        ``sync=False`` (x and out pinned) leaves the stream running so a second Engine on
        another stream can overlap its copies with this one's compute."""
        torch = self._torch
        layout = self._layout_of(x)
        x = x.contiguous()
        n = x.shape[0]
        if out is None:
            out = torch.empty((n, self.net.num_classes), dtype=torch.float32)
        if n == 0:
            return out
        nbytes = x.numel() * x.element_size()
        self._grow("_stage", nbytes, torch.uint8)
        self._grow("_logits", n * self.net.num_classes, torch.float32)
        chunk = min(int(chunk or self.chunk), n)
        ws = self._workspace(chunk)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._raise_if_out_of_range("an earlier asynchronous call")
        rc = self.lib.f8_plan_run_host(self._h, x.data_ptr(), layout, n, out.data_ptr(),
                                       self._stage.data_ptr(), self._logits.data_ptr(),
                                       ws.data_ptr(), ws.numel(), chunk, int(bool(sync)),
                                       st.cuda_stream)
        if rc == C.F8_ERR_RANGE:
            if self.range_check:
                raise InputRangeError(C.lib().f8_last_error().decode())
            return out
        C.check(rc)
        return out

    def profile(self, x, chunk=None):
        """Per-launch device times (ms) of one run_device(x): list of (op name, kind, ms)."""
        torch = self._torch
        layout = self._layout_of(x)
        n = x.shape[0]
        chunk = min(int(chunk or self.chunk), n)
        ws = self._workspace(chunk)
        out = torch.empty((n, self.net.num_classes), dtype=torch.float32, device=self.device)
        nops = len(self.plan.ops)
        ms = (ctypes.c_float * nops)()
        st = torch.cuda.current_stream(self.device)
        C.check(self.lib.f8_plan_profile(self._h, x.data_ptr(), layout, n, out.data_ptr(),
                                         ws.data_ptr(), ws.numel(), chunk, st.cuda_stream,
                                         ms, nops))
        return [(op.name, op.kind, float(ms[i])) for i, op in enumerate(self.plan.ops)]

    def kernel_names(self):
        """Kernel template that served each op in the most recent ``profile`` call."""
        out = []
        buf = ctypes.create_string_buffer(96)
        for i in range(len(self.plan.ops)):
            C.check(self.lib.f8_plan_kernel_name(self._h, i, buf, 96))
            out.append(buf.value.decode())
        return out

    def topk(self, logits, ks=(1, 5)):
        """forward_loss's prediction step (fix_train.py:698-703): indices of the max(ks) largest
        logits per image, on the logits' device."""
        return logits.topk(max(ks))[1]

    def __call__(self, x, strict=False):
        """``IntModel.forward(x)``: int32 NCHW in, float32 logits out, on x's device.
        The engine keeps the low byte of every input value where the reference's head conv consumes the
        full int32 (fix_train.py:682-692 hands it 8-bit-range integers).  Out-of-range inputs are always
        detected (InputRangeError): by this call for a CPU tensor, by the next call or
        ``check_input_range()`` for a CUDA tensor (the call is asynchronous).  ``strict`` checks up front,
        with two extra passes over x, against the symmetric range the reference's own preparation
        produces (+-127 for a signed head)."""
        torch = self._torch
        if strict and x.dtype == torch.int32:
            lo, hi = (-127, 127) if self.net.head.sym else (0, 255)
            mn, mx = int(x.min()), int(x.max())
            if mn < lo or mx > hi:
                raise ValueError(f"input range [{mn},{mx}] outside the head's [{lo},{hi}]")
        if strict and x.dtype == torch.float32 and not self.input_prep["normalize"]:
            # forward_loss asserts input >= 0 (fix_train.py:689) and (255 x).round() must stay in the
            # head's 8 bits; the normalize branch clamps to +-127 itself (fix_quant, fix_train.py:682-687)
            mn, mx = float(x.min()), float(x.max())
            if mn < 0.0 or mx > 1.0 or mn != mn:
                raise ValueError(f"float input range [{mn},{mx}] outside [0,1]: the reference asserts "
                                 f"input >= 0 and its head conv would see values beyond 8 bits")
        if x.is_cuda:
            return self.run_device(x)
        return self.run_host(x)

    forward = __call__


def compile(model_or_state_dict, arch: Optional[str] = None, head_signed: Optional[bool] = None,
            device=None, chunk: int = 256, backend=None, quant_maxpool: bool = False,
            keep_buffers: bool = False, fuse_tail=None, range_check: bool = True) -> Engine:
    """Build an Engine from a reference ``IntModel`` (module tree walked for stride / groups /
    input_symmetric), or from its ``state_dict()`` plus the architecture name -- the
    attributes the dict lacks are then re-derived from the architecture;
    ``head_signed`` mirrors FLAGS.normalize (fix_resnet.py:437-438) and defaults to False;
    ``quant_maxpool`` mirrors FLAGS.quant_maxpool (FXQMaxPool2d head pool, fix_resnet.py:331-334);
    ``range_check=False`` keeps out-of-range inputs silent (low byte used), see ``Engine.check_input_range``."""
    if hasattr(model_or_state_dict, "state_dict") and hasattr(model_or_state_dict, "head"):
        net = graph_from_module(model_or_state_dict)
        sd = model_or_state_dict.state_dict()
    else:
        sd = model_or_state_dict
        if callable(sd):            # the reference saves the bound method (fix_train.py:946)
            sd = sd()
        if arch is None:
            arch = infer_arch(sd)
        if arch not in ARCHS and not arch.startswith("resnet"):
            raise ValueError(f"unknown arch {arch!r}")
        net = graph_for(arch, bool(head_signed), quant_maxpool=bool(quant_maxpool))
    return Engine(net, sd, device=device, chunk=chunk, backend=backend, keep_buffers=keep_buffers,
                  fuse_tail=fuse_tail, range_check=range_check)
