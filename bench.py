#!/usr/bin/env python
"""bench.py -- images/sec of the int_op_only forward path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--arch resnet18] [--batch 256]
    python bench.py --impl reference ...      # the CPU implementation of the path, same line

A "step" is one pass of the hot path over one batch of synthetic images (per GPU).  At N=1
the workload is BASELINE.json configs[1]: ResNet18 int_op_only, batch 256, 224x224 synthetic
on 1xB200.  N>1 (torchrun, one rank per GPU): each rank runs its own 256-image shard (weak
scaling, no data-path collective) and the logits are all-gathered once per step (NCCL).

One JSON line on stdout (rank 0).  ``value``: inputs resident in HBM as the engine-native
NHWC u8 tensor, rotating over several distinct batches so that every step's input comes from
HBM, not L2.  ``e2e``: the same metric through the reference-facing call with HOST buffers --
pinned int32 NCHW input (the reference's tensor), host-side narrowing to the engine's 8-bit
layout, H2D copy, run, D2H copy of the logits, all inside the timed region (two engines on two
streams overlap copy and compute).
``roofline``: the dominant kernel family (dense conv implicit GEMM), algorithmic bytes of its
launches / their device time measured with CUDA events around every launch.  ``cpu_baseline``:
the CPU oracle port (oracle/) on the host cores, a bounded sample of the same workload.
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

METRIC = "images/sec (int_op_only, bit-exact)"
UNIT = "images/s"


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
    return ap.parse_args()


def config_for(args, extra=None):
    names = {"resnet18": "ResNet18", "resnet50": "ResNet50", "mobilenet_v1": "MobileNet V1",
             "mobilenet_v2": "MobileNet V2"}
    cfg = {"workload": f"{names.get(args.arch, args.arch)} int_op_only, batch={args.batch}/GPU, "
                       f"224x224 synthetic (weights seed 1234, inputs seed 1995)",
           "arch": args.arch, "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
           "parallelism": f"batch-sharded x{args.gpus}, one all-gather of logits"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's int_op_only CPU path (oracle/ is test
# infrastructure; bench.py may execute it only here, as the baseline being reported)
# ------------------------------------------------------------------------------------------
def cpu_forward_timer(arch):
    import numpy as np
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


def cpu_baseline(arch, budget_s=15.0):
    run, threads = cpu_forward_timer(arch)
    run(1)                                   # warm-up (library load, page-in)
    t1, _ = run(2)
    n = int(max(2, min(256, budget_s / max(t1 / 2, 1e-4))))
    best = min(run(n)[0] for _ in range(2))
    return {"value": n / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} images of the same workload through the C oracle port "
                      f"(oracle/f8_oracle.c + oracle/nets.py), OpenMP over {threads} threads, "
                      f"best of 2"}


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
    sample = (f"{n} images per step (bounded sample of the {args.batch}-image batch), C oracle "
              f"port of the reference CPU path, {threads} OpenMP threads")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": config_for(args, {"sample_images_per_step": n}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import f8net_b200
    from f8net_b200 import _capi as C
    from f8net_b200 import synth
    from f8net_b200.roofline import network_work, op_work
    from f8net_b200.sharded import ShardedRunner

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

    arch, B = args.arch, args.batch
    hs = synth.HEAD_SIGNED.get(arch, False)
    sd = synth.make_state_dict(arch, hs)
    kw = {}
    if args.chunk:
        kw["chunk"] = args.chunk
    if args.backend >= 0:
        kw["backend"] = args.backend
    eng = f8net_b200.compile(sd, arch=arch, head_signed=hs, device=device, **kw)
    ops_img, bytes_img, wbytes = network_work(eng.net)

    # R distinct resident input batches (engine-native NHWC u8/s8), more than L2 (126 MB)
    S = eng.net.image_size
    in_bytes = B * S * S * 4
    R = max(2, -(-160_000_000 // in_bytes))
    g = torch.Generator(device=device).manual_seed(1995 + rank)
    if hs:
        xs = [torch.randint(-127, 128, (B, S, S, 4), dtype=torch.int8, device=device, generator=g)
              for _ in range(R)]
    else:
        xs = [torch.randint(0, 256, (B, S, S, 4), dtype=torch.uint8, device=device, generator=g)
              for _ in range(R)]
    for x in xs:
        x[..., 3] = 0
    runner = ShardedRunner(eng.run_device, eng.net.num_classes) if world > 1 else None

    def step(i):
        x = xs[i % R]
        return runner(x) if runner is not None else eng.run_device(x, out=logits)

    logits = torch.empty((B, eng.net.num_classes), dtype=torch.float32, device=device)

    # CUDA graphs of the R step variants (launch-bound inner loop).  N>1: the graph holds this
    # rank's forward pass, writing into its slice of the gather buffer; the NCCL all-gather of the
    # logits follows it every step
    graphs = None
    full = mine = None
    if runner is not None:
        full = runner.gather_buffer(B, xs[0])
        mine = full[rank]
    if not args.no_graph:
        for i in range(min(R, 2)):
            step(i)                       # warm every kernel (cudaFuncSetAttribute) before capture
        torch.cuda.synchronize()
        graphs = []
        cs = torch.cuda.Stream(device)
        with torch.cuda.stream(cs):
            for i in range(R):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=cs):
                    eng.run_device(xs[i], out=logits if mine is None else mine, stream=cs)
                graphs.append(gr)
        torch.cuda.synchronize()

    def do_step(i):
        if graphs is not None:
            graphs[i % R].replay()
            if runner is not None:
                dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1))
        else:
            step(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        do_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            do_step(i)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)
    launches = eng.launches(B, C.F8_IN_NHWC4_8) * args.steps

    # ---- e2e: reference-facing call, host buffers, copies in the timed region ----
    engines = [eng, f8net_b200.compile(sd, arch=arch, head_signed=hs, device=device, **kw)]
    streams = [torch.cuda.Stream(device), torch.cuda.Stream(device)]
    lo, hi = (-127, 128) if hs else (0, 256)
    hx = [torch.randint(lo, hi, (B, 3, S, S), dtype=torch.int32).pin_memory() for _ in range(2)]
    hy = [torch.empty((B, eng.net.num_classes), dtype=torch.float32).pin_memory() for _ in range(2)]

    def e2e_steps(k):
        for i in range(k):
            j = i % 2
            streams[j].synchronize()      # the previous use of this slot's host buffers is done
            engines[j].run_host(hx[j], out=hy[j], sync=False, stream=streams[j])
        for s in streams:
            s.synchronize()

    e2e_steps(3)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    barrier()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    # run_host repacks the int32 tensor to NHWC4 bytes with the host cores (inside the timed region)
    # and ships those; F8_HOST_PACK_THREADS=0 ships the int32 tensor itself
    host_pack = os.environ.get("F8_HOST_PACK_THREADS", "") != "0"
    e2e = {"value": world * B * args.steps / e2e_dt, "unit": UNIT,
           "h2d_bytes_per_step": B * S * S * 4 if host_pack else B * 3 * S * S * 4,
           "host_tensor_bytes_per_step": B * 3 * S * S * 4,
           "d2h_bytes_per_step": B * eng.net.num_classes * 4,
           "input": "pinned int32 NCHW (the reference's tensor), two engines on two streams"
                    + ("; run_host narrows it to NHWC4 bytes on the host cores (timed) before the copy" if host_pack else ""),
           "timer": "host perf_counter around the loop, stream syncs inside"}

    # ---- the same call fed with decoded uint8 pixels [B,H,W,3]: ToTensor +
    # Normalize + forward_loss's integerisation run on the device, 4x fewer PCIe bytes ----
    hp = [torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]

    def e2e_u8_steps(k):
        for i in range(k):
            j = i % 2
            streams[j].synchronize()
            engines[j].run_host(hp[j], out=hy[j], sync=False, stream=streams[j])
        for s in streams:
            s.synchronize()

    e2e_u8_steps(3)
    barrier()
    t0 = time.perf_counter()
    e2e_u8_steps(args.steps)
    barrier()
    u8_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([u8_dt], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u8_dt = float(t.item())
    e2e_u8 = {"value": world * B * args.steps / u8_dt, "unit": UNIT,
              "h2d_bytes_per_step": B * 3 * S * S, "d2h_bytes_per_step": B * eng.net.num_classes * 4,
              "input": "pinned uint8 [B,H,W,3] decoded pixels; ToTensor + Normalize + integerisation "
                       "(fix_train.py:299-318, :676-692) on the device"}

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel family: per-launch CUDA events ----
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"
        work = op_work(eng.plan)
        acc = [0.0] * len(work)
        reps = min(args.steps, 10)
        for i in range(reps):
            for j, (_, _, t_ms) in enumerate(eng.profile(xs[i % R])):
                acc[j] += t_ms / reps
        fam = {}
        for w, t_ms in zip(work, acc):
            f = fam.setdefault(w["kind"], {"ms": 0.0, "bytes": 0.0, "ops": 0.0, "launches": 0})
            f["ms"] += t_ms
            f["bytes"] += w["bytes_per_image"] * B + w["weight_bytes"]
            f["ops"] += w["ops"] * B
            f["launches"] += 1
        total_ms = sum(acc)
        dom_kind = max(fam, key=lambda k: fam[k]["ms"])
        d = fam[dom_kind]
        kind_names = {C.F8_OP_CONVERT_INPUT: "convert_input", C.F8_OP_CONV_DENSE: "conv_dense",
                      C.F8_OP_CONV_DW: "conv_dw3x3", C.F8_OP_MAXPOOL: "maxpool",
                      C.F8_OP_POOL_REQUANT: "pool_requant", C.F8_OP_HEAD_POOL: "head_conv_pool",
                      C.F8_OP_POOL_FC: "pool_fc"}
        achieved = d["bytes"] / (d["ms"] / 1e3) / 1e9 if d["ms"] > 0 else 0.0
        # DRAM bytes per launch of this kernel family from the committed `ncu --set full` capture
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(arch, {})
            if tr.get("batch") == B and kind_names[dom_kind] in tr:
                traffic = tr[kind_names[dom_kind]]["dram_bytes_per_launch"]
        except (OSError, ValueError):
            pass
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": traffic,
            "kernel": kind_names[dom_kind], "launches_per_step": d["launches"],
            "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
            "avg_launch_ms": d["ms"] / d["launches"],
            "share_of_step": d["ms"] / total_ms if total_ms else None,
            "peak_source": peak_src,
            "int8_tops_achieved": d["ops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0,
            "method": f"CUDA event pair around every launch on the launch stream "
                      f"(f8_plan_profile), mean of {reps} steps, batch {B}",
            "whole_net": {"algorithmic_gbs": (bytes_img * B + wbytes) * world / (ms / args.steps / 1e3) / 1e9,
                          "int8_tops": ops_img * B * world / (ms / args.steps / 1e3) / 1e12,
                          "frac_of_hbm_peak": (bytes_img * B + wbytes) / (ms / args.steps / 1e3) / 1e9 / hbm_peak},
            "per_kernel_ms": {kind_names[k]: round(v["ms"], 4) for k, v in fam.items()},
        }
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(arch)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic",
            "config": config_for(args, {
                "chunk": eng.chunk, "backend": "tcgen05" if eng.backend == 1 else "mma.sync",
                "cuda_graph": graphs is not None,
                "l2": f"{R} distinct resident input batches ({R * in_bytes / 1e6:.0f} MB > 126 MB L2) rotated per step",
                "resident_input": "NHWC 8-bit [B,224,224,4]"}),
            "e2e": e2e, "e2e_uint8_input": e2e_u8, "gpu_launches": launches, "clocks": clocks.summary(),
            "roofline": roofline, "cpu_baseline": cb,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
