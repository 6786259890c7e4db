#!/usr/bin/env python
"""bench.py -- images/sec of the int_op_only forward path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--arch resnet18] [--batch 256]
    python bench.py --impl reference ...      # the CPU implementation of the path, same line

A "step" is one pass of the hot path over one batch of synthetic images (per GPU).  At N=1
the workload is BASELINE.json configs[1]: ResNet18 int_op_only, batch 256, 224x224 synthetic
on 1xB200.  N>1 (torchrun, one rank per GPU): each rank runs its own 256-image shard (weak
scaling, no data-path collective) and the logits are all-gathered once per step (NCCL); the
gather of step i overlaps the forward pass of step i+1 (two gather buffers).

One JSON line on stdout (rank 0).  ``value``: inputs resident in HBM as the engine-native
NHWC u8 tensor, rotating over several distinct batches so that every step's input comes from
HBM, not L2.  ``e2e``: the same metric through the reference-facing call with HOST buffers --
pinned int32 NCHW input (the reference's tensor), host-side narrowing to the engine's 8-bit
layout, H2D copy, run, D2H copy of the logits, all inside the timed region (two engines on two
streams overlap copy and compute).
``roofline``: the dominant kernel TEMPLATE (``top_kernels`` lists the three largest), algorithmic
bytes of its launches / their device time measured with CUDA events around every launch.
``configs``: the other BASELINE.json networks at the same per-GPU batch (MobileNet V2, ResNet50,
MobileNet V1; 256 per GPU, i.e. 2048 over 8 GPUs), each timed the same way.
``parity``: before anything is timed, a seeded check batch of N x 256 images runs (a) sharded +
all-gathered and (b) whole on every rank; the gathered logits must equal every rank's own
single-GPU result bit for bit, and the first two images are the committed golden fixture whose
logits the unmodified reference produced (tests/golden/<arch>_n2.npz).
``cpu_baseline``: the CPU oracle port (oracle/) on the host cores, a bounded sample of the same
workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _claim_stdout():
    """stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version banner to fd 1 under
    NCCL_DEBUG=VERSION, which the GPU boxes set): file descriptor 1 is pointed at stderr for the life of the process
    and the JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


RESULT_OUT = None


def emit(line):
    out = RESULT_OUT if RESULT_OUT is not None else sys.stdout
    print(json.dumps(line), file=out, flush=True)

METRIC = "images/sec (int_op_only, bit-exact)"
UNIT = "images/s"
NAMES = {"resnet18": "ResNet18", "resnet50": "ResNet50", "mobilenet_v1": "MobileNet V1",
         "mobilenet_v2": "MobileNet V2"}
OTHER_CONFIGS = ["mobilenet_v2", "resnet50", "mobilenet_v1"]      # BASELINE.json configs[2..4]
IMAGE = 224
L2_BYTES = 126_000_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="f8net_b200", choices=["f8net_b200", "reference"])
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--chunk", type=int, default=0, help="images per pass (0 = engine default)")
    ap.add_argument("--backend", type=int, default=-1, help="-1 auto, 0 mma.sync, 1 tcgen05")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE networks")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def rotation(batch):
    """How many distinct resident input batches rotate so that each step reads its input from HBM."""
    in_bytes = batch * IMAGE * IMAGE * 4
    r = max(2, -(-160_000_000 // in_bytes))
    return r + (r & 1), in_bytes


def config_for(arch, batch, gpus):
    """The workload, identical for both arms (``--impl reference`` prints the same dict)."""
    r, in_bytes = rotation(batch)
    return {"workload": f"{NAMES.get(arch, arch)} int_op_only, batch={batch}/GPU, "
                        f"224x224 synthetic (weights seed 1234, inputs seed 1995)",
            "arch": arch, "batch_per_gpu": batch, "global_batch": batch * gpus,
            "parallelism": f"batch-sharded x{gpus}, one all-gather of logits",
            "l2": f"inputs larger than L2: {r} distinct input batches "
                  f"({r * in_bytes / 1e6:.0f} MB > {L2_BYTES / 1e6:.0f} MB L2) rotated per step"}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's int_op_only CPU path (oracle/ is test
# infrastructure; bench.py may execute it only here, as the baseline being reported)
# ------------------------------------------------------------------------------------------
def cpu_forward_timer(arch):
    from f8net_b200 import synth
    from oracle import nets, oracle as O
    hs = synth.HEAD_SIGNED.get(arch, False)
    sd = synth.make_state_dict(arch, hs)
    threads = os.cpu_count() or 1
    O.set_threads(threads)

    def run(n, seed=1995):
        x = synth.make_input(arch, n, hs, seed=seed)
        t = time.perf_counter()
        y = nets.forward(arch, sd, x, hs)
        return time.perf_counter() - t, y
    return run, O.max_threads()


PORT_NOTE = ("C port of the reference's CPU path (the reference is Python on ATen int32 kernels and "
             "cannot travel to the GPU box); the port is faster than ATen's int32 convolution "
             "(SURVEY.md: 12.9 img/s at batch 32 on 8 vCPU), so GPU/CPU ratios are conservative")


def cpu_baseline(arch, budget_s=15.0):
    run, threads = cpu_forward_timer(arch)
    run(1)                                   # warm-up (library load, page-in)
    t1, _ = run(2)
    n = int(max(2, min(256, budget_s / max(t1 / 2, 1e-4))))
    best = min(run(n)[0] for _ in range(2))
    return {"value": n / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} images of the same workload through the C oracle port "
                      f"(oracle/f8_oracle.c + oracle/nets.py), OpenMP over {threads} threads, "
                      f"best of 2", "note": PORT_NOTE}


def reference_arm(args):
    """--impl reference: the CPU implementation of the path with every host thread.  The
    reference itself is pure Python on torch ATen CPU kernels (no native sources to compile
    into oracle/_ref, and /root/reference does not exist on the GPU box), so this arm times
    the C oracle port that tests/golden pins against the unmodified reference."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, threads = cpu_forward_timer(args.arch)
    run(1)
    t1, _ = run(2)
    per_img = t1 / 2
    total = args.steps + args.warmup
    n = int(max(1, min(args.batch, 90.0 / max(per_img * total, 1e-6))))
    for i in range(args.warmup):
        run(n, seed=100 + i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        run(n, seed=200 + i)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = (f"sampled: {n} images per step (bounded sample of the {args.batch}-image batch), C oracle "
              f"port of the reference CPU path, {threads} OpenMP threads")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": config_for(args.arch, args.batch, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": sample, "note": PORT_NOTE},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ArchRun:
    """One network on this rank's GPU: engine, resident rotating inputs, CUDA graphs, and (N>1)
    the overlapped all-gather of the logits."""

    def __init__(self, arch, args, device, rank, world):
        import numpy as np
        import torch

        import f8net_b200
        from f8net_b200 import synth
        from f8net_b200.roofline import network_work
        self.torch, self.np = torch, np
        self.arch, self.args, self.device, self.rank, self.world = arch, args, device, rank, world
        self.B = B = args.batch
        self.hs = hs = synth.HEAD_SIGNED.get(arch, False)
        self.sd = synth.make_state_dict(arch, hs)
        self.kw = {}
        if args.chunk:
            self.kw["chunk"] = args.chunk
        if args.backend >= 0:
            self.kw["backend"] = args.backend
        self.eng = f8net_b200.compile(self.sd, arch=arch, head_signed=hs, device=device, **self.kw)
        self.ops_img, self.bytes_img, self.wbytes = network_work(self.eng.net)
        S = self.eng.net.image_size
        self.R, self.in_bytes = rotation(B)
        g = torch.Generator(device=device).manual_seed(1995 + rank)
        self.xs = [self._random_batch(B, g) for _ in range(self.R)]
        self.classes = self.eng.net.num_classes
        self.logits = torch.empty((B, self.classes), dtype=torch.float32, device=device)
        self.og = None
        if world > 1:
            from f8net_b200.sharded import OverlappedGather
            self.og = OverlappedGather(B, self.classes, self.xs[0])
        self.graphs = None
        self.S = S

    def _random_batch(self, n, g):
        torch = self.torch
        S = IMAGE
        if self.hs:
            x = torch.randint(-127, 128, (n, S, S, 4), dtype=torch.int8, device=self.device, generator=g)
        else:
            x = torch.randint(0, 256, (n, S, S, 4), dtype=torch.uint8, device=self.device, generator=g)
        x[..., 3] = 0
        return x

    # -- parity: gathered logits == every rank's own single-GPU logits; golden fixture -----
    def parity(self):
        import torch.distributed as dist

        from f8net_b200 import synth
        from f8net_b200.sharded import ShardedRunner
        torch, np = self.torch, self.np
        B, world, rank = self.B, self.world, self.rank
        g = torch.Generator(device=self.device).manual_seed(20260)     # same batch on every rank
        xchk = self._random_batch(world * B, g)
        gx = synth.make_input(self.arch, 2, self.hs)                   # the golden fixture's two images
        g4 = np.zeros((2, IMAGE, IMAGE, 4), dtype=np.int8 if self.hs else np.uint8)
        g4[..., :3] = gx.transpose(0, 2, 3, 1)
        xchk[:2] = torch.from_numpy(g4).to(self.device)
        whole = self.eng.run_device(xchk)                              # [world*B, classes] on this GPU alone
        out = {"check_batch": world * B}
        if world > 1:
            runner = ShardedRunner(self.eng.run_device, self.classes)
            gathered = runner(xchk[rank * B:(rank + 1) * B].contiguous())
            ok = torch.tensor([int(torch.equal(gathered, whole))], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            out["gathered_logits_exact"] = bool(ok.item())
            first = gathered[:2]
        else:
            first = whole[:2]
        gold_path = os.path.join(ROOT, "tests", "golden", f"{self.arch}_n2.npz")
        if os.path.exists(gold_path):
            gold = np.load(gold_path)["logits"].astype(np.int64)
            out["golden_logits_exact"] = bool(np.array_equal(first.cpu().numpy().astype(np.int64), gold))
        torch.cuda.synchronize()
        del xchk, whole
        bad = [k for k, v in out.items() if v is False]
        if bad:
            raise SystemExit(f"bench.py parity check failed for {self.arch}: {bad} (rank {rank})")
        return out

    # -- timing of the device-resident path -----------------------------------------------
    def capture(self):
        torch = self.torch
        if self.args.no_graph:
            return
        for i in range(2):
            self.eng.run_device(self.xs[i], out=self.logits)      # warm every kernel before capture
        torch.cuda.synchronize()
        self.graphs = []
        cs = torch.cuda.Stream(self.device)
        with torch.cuda.stream(cs):
            for i in range(self.R):
                gr = torch.cuda.CUDAGraph()
                out = self.logits if self.og is None else self.og.slot(i)      # R is even: slot(i) == slot(i % R)
                with torch.cuda.graph(gr, stream=cs):
                    self.eng.run_device(self.xs[i], out=out, stream=cs)
                self.graphs.append(gr)
        torch.cuda.synchronize()

    def do_step(self, i):
        og = self.og
        if og is not None:
            og.ready(i)                       # step i-2's gather has released this buffer
        if self.graphs is not None:
            self.graphs[i % self.R].replay()
        else:
            self.eng.run_device(self.xs[i % self.R], out=self.logits if og is None else og.slot(i))
        if og is not None:
            og.submit(i)                      # async all-gather behind the forward pass just enqueued

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def time_steps(self, steps, warmup, clocks_index=None):
        import torch.distributed as dist
        torch = self.torch
        for i in range(max(warmup, 3)):
            self.do_step(i)
        if self.og is not None:
            self.og.drain()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(clocks_index) if clocks_index is not None else None
        if sampler:
            sampler.__enter__()
        self.barrier()
        e0.record()
        for i in range(steps):
            self.do_step(i)
        if self.og is not None:
            self.og.drain()                   # the current stream waits for the last two gathers
        e1.record()
        self.barrier()
        if sampler:
            sampler.__exit__()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (sampler.summary() if sampler else None)

    def summary(self, ms, steps, hbm_peak):
        from f8net_b200 import _capi as C
        per = ms / steps / 1e3
        alg = (self.bytes_img * self.B + self.wbytes) / per / 1e9        # per GPU
        return {"value": self.world * self.B * steps / (ms / 1e3), "ms_per_step": ms / steps,
                "steps": steps,
                "whole_net": {"algorithmic_gbs_per_gpu": alg, "frac_of_hbm_peak": alg / hbm_peak,
                              "int8_tops_per_gpu": self.ops_img * self.B / per / 1e12,
                              "algorithmic_mb_per_image": self.bytes_img / 1e6},
                "gpu_launches_per_step": self.eng.launches(self.B, C.F8_IN_NHWC4_8)}

    def close(self):
        self.graphs = None
        self.xs = None
        self.eng.close()
        self.torch.cuda.empty_cache()


def e2e_measure(run, args, device, barrier, world):
    """The reference-facing call with host buffers inside the timed region."""
    import torch
    import torch.distributed as dist

    import f8net_b200
    B, S = args.batch, IMAGE
    eng = run.eng
    engines = [eng, f8net_b200.compile(run.sd, arch=run.arch, head_signed=run.hs, device=device, **run.kw)]
    streams = [torch.cuda.Stream(device), torch.cuda.Stream(device)]
    lo, hi = (-127, 128) if run.hs else (0, 256)
    hy = [torch.empty((B, run.classes), dtype=torch.float32).pin_memory() for _ in range(2)]

    def timed(bufs):
        def steps(k):
            for i in range(k):
                j = i % 2
                streams[j].synchronize()      # the previous use of this slot's host buffers is done
                engines[j].run_host(bufs[j], out=hy[j], sync=False, stream=streams[j])
            for s in streams:
                s.synchronize()
        steps(3)
        barrier()
        t0 = time.perf_counter()
        steps(args.steps)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * B * args.steps / dt

    hx = [torch.randint(lo, hi, (B, 3, S, S), dtype=torch.int32).pin_memory() for _ in range(2)]
    v = timed(hx)
    raw_images = [int(e.lib.f8_plan_last_raw_images(e._h)) for e in engines]
    del hx
    import ctypes
    nthreads = ctypes.c_int(0)
    isa = eng.lib.f8_host_pack_info(ctypes.byref(nthreads)).decode()
    # run_host repacks the int32 tensor to NHWC4 bytes with the host cores (inside the timed region)
    # and ships those; F8_HOST_PACK_THREADS=0 ships the int32 tensor itself
    host_pack = os.environ.get("F8_HOST_PACK_THREADS", "") != "0"
    e2e = {"value": v, "unit": UNIT,
           "h2d_bytes_per_step": (B * S * S * 4 + 2 * raw_images[0] * S * S * 4) if host_pack else B * 3 * S * S * 4,
           "host_tensor_bytes_per_step": B * 3 * S * S * 4,
           "d2h_bytes_per_step": B * run.classes * 4,
           "input": "pinned int32 NCHW (the reference's tensor), two engines on two streams"
                    + ("; run_host splits the batch in 16-image sub-batches: the host cores narrow them to NHWC4 bytes from "
                       "the front (timed), the copy engine ships raw ones from the back, narrowed on the device" if host_pack else ""),
           "host_narrowing": {"simd": isa, "threads_per_rank": nthreads.value,
                              "images_shipped_raw_last_step": raw_images, "of": B},
           "host_bound": "the int32 tensor is 602 KB per image that either a host core or the copy engine must read "
                         "from host DRAM: this number is bound by host memory bandwidth and PCIe, not by the GPU",
           "timer": "host perf_counter around the loop, stream syncs inside"}
    # the same call fed with decoded uint8 pixels [B,H,W,3]: ToTensor +
    # Normalize + forward_loss's integerisation run on the device, 4x fewer PCIe bytes
    hp = [torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    v8 = timed(hp)
    e2e_u8 = {"value": v8, "unit": UNIT,
              "h2d_bytes_per_step": B * 3 * S * S, "d2h_bytes_per_step": B * run.classes * 4,
              "input": "pinned uint8 [B,H,W,3] decoded pixels; ToTensor + Normalize + integerisation "
                       "(fix_train.py:299-318, :676-692) on the device"}
    engines[1].close()
    return e2e, e2e_u8


def roofline_for(run, args, hbm_peak, peak_src, ms_step, int8_peak):
    """Per kernel template: algorithmic bytes / device time of its launches (live CUDA events)."""
    from f8net_b200.roofline import op_work
    eng, B = run.eng, run.B
    work = op_work(eng.plan)
    acc = [0.0] * len(work)
    reps = min(args.steps, 10)
    for i in range(reps):
        for j, (_, _, t_ms) in enumerate(eng.profile(run.xs[i % run.R])):
            acc[j] += t_ms / reps
    names = eng.kernel_names()
    tmpl = {}
    for w, t_ms, name in zip(work, acc, names):
        if not name:
            continue
        f = tmpl.setdefault(name, {"ms": 0.0, "bytes": 0.0, "ops": 0.0, "launches": 0})
        f["ms"] += t_ms
        f["bytes"] += w["bytes_per_image"] * B + w["weight_bytes"]
        f["ops"] += w["ops"] * B
        f["launches"] += 1
    total_ms = sum(acc)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(run.arch, {})
        if traffic.get("batch") != B:
            traffic = {}
    except (OSError, ValueError):
        pass

    def entry(name):
        d = tmpl[name]
        ach = d["bytes"] / (d["ms"] / 1e3) / 1e9 if d["ms"] > 0 else 0.0
        tops = d["ops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0
        tr = traffic.get(name, {}).get("dram_bytes_per_launch")
        return {"kernel": name, "launches_per_step": d["launches"],
                "achieved": ach, "frac": ach / hbm_peak,
                "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
                "avg_launch_ms": d["ms"] / d["launches"],
                "share_of_step": d["ms"] / total_ms if total_ms else None,
                "int8_tops_achieved": tops,
                "int8_frac": tops / int8_peak["tops"] if int8_peak else None,
                "traffic": tr}

    order = sorted(tmpl, key=lambda k: -tmpl[k]["ms"])
    top = [entry(k) for k in order[:3]]
    dom = top[0]
    roofline = {
        "bound": "hbm", "achieved": dom["achieved"], "peak": hbm_peak, "unit": "GB/s",
        "frac": dom["frac"], "traffic": dom["traffic"],
        "kernel": dom["kernel"], "launches_per_step": dom["launches_per_step"],
        "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
        "avg_launch_ms": dom["avg_launch_ms"], "share_of_step": dom["share_of_step"],
        "peak_source": peak_src,
        "int8_tops_achieved": dom["int8_tops_achieved"],
        "int8_peak_tops": int8_peak["tops"] if int8_peak else None,
        "int8_peak_source": int8_peak["shape"] if int8_peak else None,
        "int8_frac": dom["int8_frac"],
        "method": f"CUDA event pair around every launch on the launch stream (f8_plan_profile), mean of "
                  f"{reps} steps, batch {B}, launches grouped by kernel template (f8_plan_kernel_name); the "
                  f"sum of event-bracketed launches is {total_ms:.3f} ms against {ms_step:.3f} ms for the "
                  f"graph-replayed step, so per-launch fractions are understated by up to "
                  f"{100 * max(0.0, total_ms / ms_step - 1):.0f} %",
        "traffic_source": traffic.get("source"),
        "top_kernels": top,
        "per_kernel_ms": {k: round(tmpl[k]["ms"], 4) for k in order},
    }
    return roofline


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    from f8net_b200 import _capi as C
    from f8net_b200.peaks import measure_int8_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"

    # ---- headline network ----
    run = ArchRun(args.arch, args, device, rank, world)
    parity = run.parity()
    run.capture()
    ms, clocks = run.time_steps(args.steps, args.warmup, clocks_index=local)
    head = run.summary(ms, args.steps, hbm_peak)
    launches = head["gpu_launches_per_step"] * args.steps

    e2e = e2e_u8 = None
    if not args.no_e2e:
        e2e, e2e_u8 = e2e_measure(run, args, device, run.barrier, world)

    roofline = cb = None
    if rank == 0:
        int8_peak = None
        try:
            int8_peak = measure_int8_peak(device)
        except C.F8Error as e:                       # e.g. mma.sync-only device
            int8_peak = None
            print(f"int8 peak measurement unavailable: {e}", file=sys.stderr)
        roofline = roofline_for(run, args, hbm_peak, peak_src, ms / args.steps, int8_peak)
        roofline["whole_net"] = dict(head["whole_net"], int8_frac=(
            head["whole_net"]["int8_tops_per_gpu"] / int8_peak["tops"] if int8_peak else None))
    engine_info = {"chunk": run.eng.chunk, "backend": "tcgen05" if run.eng.backend == 1 else "mma.sync",
                   "cuda_graph": run.graphs is not None, "resident_input": "NHWC 8-bit [B,224,224,4]",
                   "gather": "async all-gather of step i overlaps the forward pass of step i+1" if world > 1 else None}
    run.close()
    if world > 1:
        dist.barrier()

    # ---- the other BASELINE.json networks at the same per-GPU batch ----
    configs = []
    if not args.no_configs:
        steps2 = max(5, min(args.steps, 50))
        for arch in OTHER_CONFIGS:
            if arch == args.arch:
                continue
            r2 = ArchRun(arch, args, device, rank, world)
            p2 = r2.parity()
            r2.capture()
            ms2, _ = r2.time_steps(steps2, min(args.warmup, 5))
            s2 = r2.summary(ms2, steps2, hbm_peak)
            configs.append(dict(config=config_for(arch, args.batch, world), parity=p2, **s2))
            launches += s2["gpu_launches_per_step"] * steps2
            r2.close()
            if world > 1:
                dist.barrier()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args.arch)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": config_for(args.arch, args.batch, world),
            "engine": engine_info, "parity": parity,
            "e2e": e2e, "e2e_uint8_input": e2e_u8, "gpu_launches": launches, "clocks": clocks,
            "roofline": roofline, "configs": configs, "cpu_baseline": cb,
        }
        emit(line)


def main():
    global RESULT_OUT
    args = parse()
    RESULT_OUT = _claim_stdout()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
