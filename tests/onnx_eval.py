"""Re-executes an ONNX graph written by f8net_b200.onnx_export (decoded with its own reader) with
numpy, int32 two's-complement wrap everywhere; Conv / Gemm go through the CPU oracle's integer
kernels (oracle/ -- test infrastructure).  Only the operator set the exporter emits."""
import numpy as np

from oracle import oracle as O


def _i32(a):
    return np.asarray(a).astype(np.int64).astype(np.int32) if np.asarray(a).dtype != np.int32 else np.asarray(a)


def _wrap(a64):
    return (a64 & 0xFFFFFFFF).astype(np.uint32).view(np.int32) if a64.ndim else np.int32(np.uint32(int(a64) & 0xFFFFFFFF))


def run(model, x):
    env = dict(model["initializers"])
    env["input"] = np.ascontiguousarray(x, dtype=np.int32)
    for node in model["nodes"]:
        op, a = node["op"], node["attrs"]
        i = [env[n] for n in node["inputs"]]
        if op == "Conv":
            assert a["dilations"] == [1, 1] and len(set(a["pads"])) == 1 and len(set(a["strides"])) == 1
            assert list(i[1].shape[2:]) == a["kernel_shape"]
            y = O.conv2d(_i32(i[0]), i[1], i[2], a["strides"][0], a["pads"][0], a["group"])
        elif op == "Gemm":
            assert a["transB"] == 1 and a["alpha"] == 1.0 and a["beta"] == 1.0
            y = O.linear(_i32(i[0]), i[1], i[2])[0]
        elif op == "Relu":
            y = np.maximum(i[0], 0)
        elif op in ("Add", "Sub", "Mul"):
            f = {"Add": np.add, "Sub": np.subtract, "Mul": np.multiply}[op]
            assert i[0].dtype == np.int32 and i[1].dtype == np.int32
            y = _wrap(f(i[0].astype(np.int64), i[1].astype(np.int64)))
        elif op == "Div":
            q = np.abs(i[0].astype(np.int64)) // np.abs(i[1].astype(np.int64))       # ONNX integer Div truncates
            y = (q * np.sign(i[0].astype(np.int64)) * np.sign(i[1].astype(np.int64))).astype(np.int32)
        elif op == "Mod":
            assert a.get("fmod", 0) == 0
            y = np.mod(i[0], i[1])            # sign of the divisor (ONNX Mod, fmod = 0) == torch.remainder
        elif op == "Equal":
            y = i[0] == i[1]
        elif op == "Where":
            y = np.where(i[0], i[1], i[2])
        elif op == "Max":
            y = np.maximum(i[0], i[1])
        elif op == "Min":
            y = np.minimum(i[0], i[1])
        elif op == "Cast":
            if a["to"] == 1:
                y = i[0].astype(np.float32)
            elif a["to"] == 7:
                y = i[0].astype(np.int64)
            elif i[0].dtype == np.float32:     # float -> int32 as the reference's x86 .int(): indefinite on overflow
                f = i[0]
                ok = (f >= -2147483648.0) & (f < 2147483648.0)
                y = np.where(ok, np.where(ok, f, 0).astype(np.int64), -(1 << 31)).astype(np.int32)
            else:
                y = _wrap(i[0].astype(np.int64))
        elif op == "MaxPool":
            assert a["kernel_shape"] == [3, 3] and a["strides"] == [2, 2] and a["pads"] == [1, 1, 1, 1]
            v = i[0]
            n, c, h, w = v.shape
            lowest = -np.inf if v.dtype == np.float32 else np.iinfo(v.dtype).min
            p = np.full((n, c, h + 2, w + 2), lowest, dtype=v.dtype)
            p[:, :, 1:-1, 1:-1] = v
            ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
            y = np.full((n, c, ho, wo), lowest, dtype=v.dtype)
            for r in range(3):
                for s in range(3):
                    y = np.maximum(y, p[:, :, r:r + 2 * ho:2, s:s + 2 * wo:2])
        elif op == "ReduceSum":
            y = i[0].sum(axis=tuple(a["axes"]), keepdims=bool(a["keepdims"]))
        elif op == "Identity":
            y = i[0]
        else:
            raise NotImplementedError(op)
        env[node["outputs"][0]] = y
    return env["output"]
