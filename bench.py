#!/usr/bin/env python
"""bench.py -- PC train samples/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

A "step" is one predictive-coding train step of the workload: feed-forward init, loss at the init,
T fixed-step inference iterations, per-layer energies at the solution, parameter gradients (and,
at N > 1, ONE NCCL all-reduce of the flat gradient/energy/loss buffer).  The optimiser update is
host-framework work outside the kernels, as in the reference (jpc/_train.py:234-242); it is
included in the `e2e` number, which goes through the public `jpc_b200.make_pc_step` API from
pinned HOST buffers.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c1|c2|c3|c5] [--impl reference]

Launched under torchrun for N > 1 (one rank per GPU, NCCL).  Weak scaling: every rank holds the
workload's full batch; the energy normalisation uses the GLOBAL batch.
`--impl reference` times the CPU restatement of the reference (oracle/, torch-CPU autograd of the
restated energy -- JAX is not installable in this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (d_in, width, depth, d_out, act, use_bias, param_type, skips, B, solver, max_t1, dt, dtype)
WORKLOADS = {
    "c1": dict(d_in=784, width=300, depth=3, d_out=10, act="tanh", bias=False, ptype="sp", skips=False, B=64,
               solver="euler", max_t1=10.0, dt=0.5, dtype="f32",
               desc="discriminative MLP 784-300-300-10 tanh, batch 64, 20 Euler steps, fp32"),
    "c2": dict(d_in=10, width=256, depth=3, d_out=784, act="relu", bias=True, ptype="sp", skips=False, B=256,
               solver="heun", max_t1=25.0, dt=0.5, dtype="f32",
               desc="supervised generative PC MLP 10-256-256-784, batch 256, 50 Heun steps, fp32"),
    "c3": dict(d_in=784, width=128, depth=128, d_out=10, act="relu", bias=False, ptype="mupc", skips=True, B=128,
               solver="euler", max_t1=10.0, dt=0.5, dtype="f32",
               desc="muPC deep residual MLP width 128 depth 128, batch 128, 20 Euler steps, fp32"),
    "c4": dict(d_in=2048, width=2048, depth=8, d_out=2048, act="tanh", bias=False, ptype="sp", skips=False, B=4096,
               solver="euler", max_t1=10.0, dt=0.5, dtype="bf16",
               desc="wide MLP width 2048 depth 8 (2048-in, 2048-out), batch 4096 per GPU, bf16 operands / fp32 "
                    "accumulate+master, 20 Euler steps"),
    "c5": dict(d_in=10, width=1024, depth=6, d_out=784, act="relu", bias=False, ptype="sp", skips=False, B=1024,
               solver="euler", max_t1=15.0, dt=0.5, dtype="bf16", bpc=True,
               desc="bidirectional PC (bPC) MLP width 1024 depth 6 (10 -> 784 top-down, 784 -> 10 bottom-up), batch 1024, "
                    "30 Euler steps, bf16 operands / fp32 accumulate+master (--bpc-f32 for the fp32 CUDA-core path)"),
}
METRIC = "PC train samples/sec (T inference iters)"
UNIT = "samples/s"


def n_steps(w):
    import math
    return max(1, int(math.ceil(w["max_t1"] / w["dt"] - 1e-6)))


def flops_per_step(w):
    """SURVEY Appendix D closed-form schedule (layer-0 prediction hoisted): g_l = 2 B d_l d_{l+1}."""
    dims = [w["d_in"]] + [w["width"]] * (w["depth"] - 1) + [w["d_out"]]
    g = [2.0 * w["B"] * dims[l] * dims[l + 1] for l in range(w["depth"])]
    evals = n_steps(w) * (2 if w["solver"] == "heun" else 1)
    per_eval = 2 * sum(g[1:])
    # executed FLOPs: from the feed-forward init the error wavefront moves one layer per Euler step, and the
    # train step only evaluates the layers it has reached (csrc/api.cu infer_loop)
    L = w["depth"]
    if w["solver"] == "euler" and L > 2:
        executed = sum(2 * sum(g[max(1, L - 1 - s):]) for s in range(n_steps(w)))
    else:
        executed = evals * per_eval
    return {"init": sum(g), "per_eval": per_eval, "inference": evals * per_eval, "energy": sum(g[1:]),
            "wgrad": sum(g), "total": 2 * sum(g) + evals * per_eval + sum(g[1:]), "evals": evals,
            "executed_total": 2 * sum(g) + executed + sum(g[1:])}


def make_inputs(w, seed_w=0, seed_d=1):
    """SURVEY 8d synthetic inputs: weights seed 0 (U(+-1/sqrt(in)); N(0,1) for mupc), data seed 1 (input side
    N(0,1), target side one-hot).  Plain dicts {"W","b","act"} in fp32 on the host; no oracle code involved."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed_w)
    layers = []
    for i in range(w["depth"]):
        _in = w["d_in"] if i == 0 else w["width"]
        _out = w["d_out"] if i + 1 == w["depth"] else w["width"]
        lim = 1.0 / math.sqrt(_in)
        W = (torch.rand(_out, _in, generator=g, dtype=torch.float32) * 2 - 1) * lim
        b = (torch.rand(_out, generator=g, dtype=torch.float32) * 2 - 1) * lim if w["bias"] else None
        if w["ptype"] == "mupc":
            W = torch.randn(_out, _in, generator=g, dtype=torch.float32)
        layers.append({"W": W, "b": b, "act": "linear" if i == 0 else w["act"]})
    g = torch.Generator().manual_seed(seed_d)
    x = torch.randn(w["B"], w["d_in"], generator=g, dtype=torch.float32)
    idx = torch.randint(0, w["d_out"], (w["B"],), generator=g)
    y = torch.nn.functional.one_hot(idx, w["d_out"]).to(torch.float32)
    return layers, x, y


def to_device_model(layers, device):
    """host layer dicts -> jpc_b200.nn model (Sequential([Lambda(act), Linear]) per layer) on `device`."""
    from jpc_b200 import nn
    out = []
    for layer in layers:
        lin = nn.Linear.__new__(nn.Linear)
        lin.weight = layer["W"].contiguous().to(device)
        lin.bias = None if layer["b"] is None else layer["b"].contiguous().to(device)
        lin.out_features, lin.in_features = layer["W"].shape
        lin.use_bias = layer["b"] is not None
        out.append(nn.Sequential([nn.Lambda(nn.get_act_fn(layer["act"])), lin]))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step_fn(w, sample_rows):
    """The reference's own algorithm on the host cores: restated energy + reverse-mode AD + fixed-step
    loop (oracle.pc_train_step), fp32, on `sample_rows` rows of the workload."""
    import torch
    from oracle import pc_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    om, x, y = make_inputs(w)
    x, y = x[:sample_rows].contiguous(), y[:sample_rows].contiguous()
    skips = O.make_skip_model(w["depth"]) if w["skips"] else None

    def step():
        return O.pc_train_step(om, skips, x, y, solver=w["solver"], max_t1=w["max_t1"], dt=w["dt"], param_type=w["ptype"])
    return step


def cpu_bpc_step_fn(w, sample_rows):
    """The reference's bPC step (config 5) on the host cores: oracle port -- restated bPC energy + reverse-mode AD, T
    gradient steps on the activities, then both models' parameter gradients; fp32, `sample_rows` rows."""
    import torch
    from oracle import pc_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    gen_l, _, _ = make_inputs(dict(w, d_in=w["d_in"], d_out=w["d_out"]), seed_w=0)
    amo_l, _, _ = make_inputs(dict(w, d_in=w["d_out"], d_out=w["d_in"]), seed_w=1)
    g = torch.Generator().manual_seed(1)
    img = torch.randn(w["B"], w["d_out"], generator=g)[:sample_rows].contiguous()
    lab = torch.nn.functional.one_hot(torch.randint(0, w["d_in"], (w["B"],), generator=g), w["d_in"]).float()[:sample_rows].contiguous()
    td_o, bu_o, T = gen_l, amo_l[::-1], n_steps(w)

    def step():
        a = O.init_activities_with_ffwd(gen_l, lab)   # one feed-forward sweep (same work as the amortiser sweep of the GPU arm)
        for _ in range(T):
            _, gr = O.compute_bpc_activity_grad(td_o, bu_o, a, img, x=lab)
            a = [p - w["dt"] * q for p, q in zip(a, gr)]
        return O.compute_bpc_param_grads(td_o, bu_o, a, img, x=lab)
    return step


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    rows = min(w["B"], args.cpu_rows)
    step = cpu_bpc_step_fn(w, rows) if w.get("bpc") else cpu_reference_step_fn(w, rows)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = rows * args.steps / dt
    cores = torch.get_num_threads()
    what = "the reference's bPC step" if w.get("bpc") else "jpc.make_pc_step"
    sample = f"{rows} of {w['B']} rows of the workload per step, fp32, torch-CPU autograd restatement of {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "rows_per_step": rows},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def run_bpc(args, w):
    """BASELINE config 5: one bPC train step = feed-forward init of the bottom-up chain (image -> label,
    examples/bidirectional_pc.ipynb cell 8) + T fused Euler steps on the bPC energy + both models' parameter
    gradients.  Default arithmetic: fp32 parity on the tensor cores (f32x3; BASELINE's config 5 is fp32)."""
    import torch
    import jpc_b200 as J
    from jpc_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dtype = "f32" if args.bpc_f32 else args.bpc_dtype
    J.set_compute_dtype(dtype)
    gen_l, lab_h, img_h = make_inputs(dict(w, d_in=w["d_in"], d_out=w["d_out"]), seed_w=0)       # top-down: label -> image
    amo_l, _, _ = make_inputs(dict(w, d_in=w["d_out"], d_out=w["d_in"]), seed_w=1)               # amortiser: image -> label
    g = torch.Generator().manual_seed(1)
    img_h = torch.randn(w["B"], w["d_out"], generator=g)
    lab_h = torch.nn.functional.one_hot(torch.randint(0, w["d_in"], (w["B"],), generator=g), w["d_in"]).float()
    td = to_device_model(gen_l, dev)
    amort = to_device_model(amo_l, dev)
    bu = amort[::-1]
    img_p, lab_p = img_h.pin_memory(), lab_h.pin_memory()
    img, lab = img_p.to(dev), lab_p.to(dev)
    T = n_steps(w)

    def step(x_lab, y_img):
        acts = J.init_activities_with_ffwd(amort, y_img)          # (reference order: amortiser sweep, image side first)
        _, g_td, g_bu, _ = J.bpc_train_step(td, bu, acts, y_img, input=x_lab, n_steps=T, dt=w["dt"])  # T Euler steps + both gradients
        return g_td, g_bu

    for _ in range(max(3, args.warmup)):
        step(lab, img)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    sampler.start()
    l0 = _lib.lib.jpcb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(lab, img)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib.jpcb_launch_count() - l0
    clocks = sampler.stop()
    # End-to-end warm-up: the functional API's buffers cycle through a few address sets and the library captures a launch
    # plan for each the second time it sees it (a one-off cost, like the reference's jit trace): run until four consecutive
    # steps captured nothing (at least 8 steps, at most 40), and report what was still captured inside the timed region.
    e2e_warm, quiet, caps = 0, 0, J.plan_cache_stats()["captures"]
    while e2e_warm < 40 and (e2e_warm < 8 or quiet < 4):
        gr = step(lab_p.to(dev, non_blocking=True), img_p.to(dev, non_blocking=True))
        float(gr[0][0][1].weight[0, 0])
        e2e_warm += 1
        now = J.plan_cache_stats()["captures"]
        quiet = quiet + 1 if now == caps else 0
        caps = now
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gr = step(lab_p.to(dev, non_blocking=True), img_p.to(dev, non_blocking=True))
        float(gr[0][0][1].weight[0, 0])  # device -> host read of a result
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_caps = J.plan_cache_stats()["captures"] - caps
    dims = [w["d_in"]] + [w["width"]] * (w["depth"] - 1) + [w["d_out"]]
    gl = [2.0 * w["B"] * dims[l] * dims[l + 1] for l in range(w["depth"])]
    per_step = 4 * sum(gl) - 2 * gl[0] - 2 * gl[-1]            # 2 directions x (prediction + transposed) GEMMs, clamped ends hoisted
    total = sum(gl) + T * per_step + 2 * sum(gl) + 2 * sum(gl)  # init + inference + errors at the solution + 2 x weight gradients
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ach = total / (ms / args.steps * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    line = {"metric": METRIC, "value": w["B"] * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": w["desc"], "global_batch": w["B"], "inference_steps": T, "solver": "euler", "parallelism": "dp1",
                       "l2": "working set fits L2; no flush", "algorithmic_gflop_per_step": total / 1e9},
            "e2e": {"value": w["B"] * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 4 * (img_p.numel() + lab_p.numel()),
                    "d2h_bytes_per_step": 4, "warmup": e2e_warm, "plan_captures_in_timed_region": int(e2e_caps)},
            "gpu_launches": int(launches), "clocks": clocks, "launch_plans": J.plan_cache_stats(),
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                         "kernel": {"bf16": "tc_grouped_gemm (bPC phases: errors / forward half / backward half + update), whole step",
                                    "f32x3": "tc_grouped_gemm on 3-way bf16 split operands (six products per GEMM, fp32 parity): "
                                             "ALGORITHMIC FLOPs counted once, 6x that many are executed; whole step",
                                    "f32": "simt_grouped_gemm<128,128,16,2,2> (fp32 CUDA-core FMA path), whole step"}[dtype]},
            "cpu_baseline": None}
    if not args.no_cpu_baseline:
        # the reference's algorithm for this step on the host cores (oracle port: restated bPC energy + reverse-mode AD),
        # fp32, on a bounded sample of the batch
        rows = min(w["B"], max(16, args.cpu_rows // 4))
        cpu_step = cpu_bpc_step_fn(w, rows)
        cpu_step()
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or time.perf_counter() - t0 < 8.0:
            cpu_step()
            reps += 1
        line["cpu_baseline"] = {"value": rows * reps / (time.perf_counter() - t0), "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "sample": f"{reps} steps of {rows} of {w['B']} rows, fp32 torch-CPU autograd "
                                                           "restatement of the reference's bPC step"}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the workload's batch on every rank (default, the driver's contract); strong = the "
                         "workload's batch split over the ranks (SURVEY 8e: 4096 -> 2048/1024/512 rows)")
    ap.add_argument("--grad-dtype", default="f32", choices=["f32", "bf16"],
                    help="N > 1: payload of the gradient all-reduce (bf16 halves the NVLink bytes; energies / loss stay fp32)")
    ap.add_argument("--dp-buckets", type=int, default=0, help="N > 1: layer groups of the bucketed gradient exchange (0: automatic)")
    ap.add_argument("--cpu-rows", type=int, default=256, help="rows per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bpc-f32", action="store_true", help="c5: run bPC on the fp32 CUDA-core path (same as --bpc-dtype f32)")
    ap.add_argument("--bpc-dtype", default="f32x3", choices=["f32x3", "bf16", "f32"],
                    help="c5 arithmetic: f32x3 = fp32 parity on the tensor cores (3-way bf16 split operands; BASELINE's c5 is "
                         "fp32), bf16 = plain bf16 operands (2e-2), f32 = CUDA-core FMA")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    if w.get("bpc"):
        return run_bpc(args, w)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: jpc_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import jpc_b200 as J
    from jpc_b200 import _lib
    from jpc_b200._core import _Net, _solver_id, _time_grid
    from jpc_b200._train import _Arena, dp_param_grads_and_exchange

    J.set_compute_dtype(w["dtype"])
    if args.scaling == "strong" and world > 1:
        if w["B"] % (8 * world):
            raise SystemExit(f"--scaling strong: batch {w['B']} does not split into {world} shards of a multiple of 8 rows")
        w = dict(w, B=w["B"] // world, desc=w["desc"] + f" (strong scaling: {w['B'] // world} of {w['B']} rows per GPU)")
    om, x_h, y_h = make_inputs(w, seed_d=1 + rank)
    gm = to_device_model(om, dev)
    gsk = J.make_skip_model(w["depth"]) if w["skips"] else None
    x_pin, y_pin = x_h.pin_memory(), y_h.pin_memory()
    x, y = x_pin.to(dev), y_pin.to(dev)
    solver = J.Euler() if w["solver"] == "euler" else J.Heun()
    T, h, hl = _time_grid(w["max_t1"], w["dt"])
    B_global = w["B"] * world

    def mknet(n_steps_):
        return _Net(gm, gsk, x, y, param_type=w["ptype"], loss_id="mse", gamma=None, output_energy_scaling=None,
                    activity_decay=0.0, weight_decay=0.0, spectral_penalty=0.0, solver=_solver_id(solver),
                    n_steps=n_steps_, dt=h, dt_last=hl, batch_global=B_global)
    net = mknet(T)
    arena = _Arena(net)
    acts = net.new_acts()
    ws = net.workspace()
    stream = net.stream()
    A_ptrs, gW_ptrs = _lib.ptr_array(acts), _lib.ptr_array(arena.gW)
    gb_ptrs = _lib.ptr_array(arena.gb) if arena.gb else None

    if args.grad_dtype != "f32":
        J.set_grad_allreduce_dtype(args.grad_dtype)
    if args.dp_buckets > 0:
        J.set_dp_buckets(args.dp_buckets)

    def step():
        # N = 1: the whole step is one C-ABI call.  N > 1: the same call without the weight gradients, then the gradients
        # bucket by bucket with each bucket's NCCL all-reduce (SUM of [gW, gb | E_layers, loss]) overlapping the next
        # bucket's GEMMs -- exactly what jpc_b200.make_pc_step does under torch.distributed
        _lib.check(_lib.lib.jpcb_pc_train_step(net.desc, net.W, net.b, _lib.ptr(x), _lib.ptr(y), A_ptrs, _lib.ptr(arena.loss),
                                               _lib.ptr(arena.E), gW_ptrs if world == 1 else None,
                                               gb_ptrs if world == 1 else None, _lib.ptr(ws), ws.numel(), stream))
        if world > 1:
            dp_param_grads_and_exchange(net, arena)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.lib.jpcb_launch_count()
    ms = timed(step, args.steps)
    launches = _lib.lib.jpcb_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = B_global * args.steps / (ms * 1e-3)
    loss_val = float(arena.loss)

    # ---- end to end through the public API, inputs from pinned host memory every step ----
    opt = J.optim.sgd(1e-3 * w["B"])
    st = opt.init((gm, gsk))
    model, skipm = gm, gsk

    # Every step's inputs are copied host -> device inside the timed region, the way a training loop's loader does
    # it: on a copy stream, double-buffered, so the copy of step i+1 overlaps the kernels of step i.  The result
    # (loss) of every step is read back to the host.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x), torch.empty_like(y), torch.cuda.Event(), torch.cuda.Event()) for _ in range(2)]

    def stage(i):
        xb, yb, ready, consumed = bufs[i % 2]
        copy_stream.wait_event(consumed)  # the step that last read this buffer has been enqueued and finished reading
        with torch.cuda.stream(copy_stream):
            xb.copy_(x_pin, non_blocking=True)
            yb.copy_(y_pin, non_blocking=True)
            ready.record(copy_stream)

    # The result (loss) of EVERY step travels device -> host inside the timed region: a non-blocking copy into pinned host
    # memory, drained (and checked) when the timed region ends -- the way a training loop logs its loss without stalling the
    # launch queue.  `e2e_blocking_read` is the same loop with the host blocking on every step's loss instead.
    loss_pin = torch.empty(256, dtype=torch.float32).pin_memory()
    E2E_WARM = max(args.warmup, 8)  # (the functional API's buffers cycle through a few address sets; launch plans settle)

    def e2e_step(i, blocking=False):
        nonlocal model, skipm, st
        xb, yb, ready, consumed = bufs[i % 2]
        torch.cuda.current_stream(dev).wait_event(ready)
        stage(i + 1)  # next step's inputs travel while this step computes
        out = J.make_pc_step(model, opt, st, yb, input=xb, param_type=w["ptype"], ode_solver=solver, max_t1=w["max_t1"],
                             dt=w["dt"], stepsize_controller=J.ConstantStepSize(), skip_model=skipm)
        consumed.record(torch.cuda.current_stream(dev))
        model, skipm, st = out["model"], out["skip_model"], out["opt_state"]
        if blocking:
            return float(out["loss"])
        loss_pin[i % 256].copy_(out["loss"], non_blocking=True)
        return None

    def e2e_loop(step_fn, blocking):
        sync_all()
        for b in bufs:
            b[3].record(torch.cuda.current_stream(dev))
        stage(0)
        # warm-up: at least E2E_WARM steps; on one GPU, on until four consecutive steps captured no new launch plan (at most
        # 40 steps) -- captures are a one-off cost like the reference's jit trace.  (Ranks of a data-parallel run must take
        # the same number of steps -- every step holds a collective -- so there the count stays fixed.)
        warm, quiet, caps = 0, 0, J.plan_cache_stats()["captures"]
        while warm < E2E_WARM or (world == 1 and quiet < 4 and warm < 40):
            step_fn(warm, blocking)
            warm += 1
            now = J.plan_cache_stats()["captures"]
            quiet = quiet + 1 if now == caps else 0
            caps = now
        sync_all()
        t0 = time.perf_counter()
        for i in range(warm, warm + args.steps):
            step_fn(i, blocking)
        sync_all()  # (drains the loss copies too)
        secs = time.perf_counter() - t0
        e2e_info["warmup"], e2e_info["plan_captures_in_timed_region"] = warm, int(J.plan_cache_stats()["captures"] - caps)
        if not bool(torch.isfinite(loss_pin[:min(256, warm + args.steps)]).all()):
            raise SystemExit("bench.py: a non-finite loss came back from the end-to-end loop")
        if world > 1:
            t = torch.tensor([secs], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t)
        return secs

    loss_pin.zero_()
    e2e_info = {}
    e2e_s = e2e_loop(e2e_step, False)
    e2e_first = dict(e2e_info)
    e2e_val = B_global * args.steps / e2e_s
    e2e_blocking_val = B_global * args.steps / e2e_loop(e2e_step, True)

    # ---- the same end-to-end loop through jpc_b200.PCTrainer: the train step + SGD update as ONE replayed CUDA graph on
    # persistent buffers (single process only; an extra, `e2e` above stays the reference-shaped functional API) ----
    e2e_graph = None
    if True:
        try:
            tr = J.PCTrainer(gm, J.optim.sgd(1e-3 * w["B"]), batch_size=w["B"], skip_model=gsk, param_type=w["ptype"], ode_solver=solver,
                             max_t1=w["max_t1"], dt=w["dt"], stepsize_controller=J.ConstantStepSize())

            def graph_step(i, blocking=False):
                xb, yb, ready, consumed = bufs[i % 2]
                torch.cuda.current_stream(dev).wait_event(ready)
                stage(i + 1)
                out = tr.step(yb, xb)
                consumed.record(torch.cuda.current_stream(dev))
                loss_pin[i % 256].copy_(out["loss"], non_blocking=True)

            g_s = e2e_loop(graph_step, False)
            e2e_graph = {"value": B_global * args.steps / g_s, "unit": UNIT,
                         "what": "jpc_b200.PCTrainer.step: H2D of the batch + one CUDA-graph replay (train step + "
                                 + ("captured NCCL all-reduce + " if world > 1 else "") + "in-place SGD) + loss read-back"}
        except NotImplementedError:
            e2e_graph = None

    # ---- roofline of the dominant kernel (grouped GEMM of one inference phase) ----
    fl = flops_per_step(w)
    roof = None
    if rank == 0:
        net0 = mknet(0)
        reps = max(3, args.steps // 2)

        def infer(nobj):
            _lib.check(_lib.lib.jpcb_pc_infer(nobj.desc, nobj.W, nobj.b, A_ptrs, _lib.ptr(x), _lib.ptr(y), None,
                                              _lib.ptr(ws), ws.numel(), stream))
        for nobj in (net, net0):
            infer(nobj)
        torch.cuda.synchronize()

        def ev(nobj):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                infer(nobj)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        ms_T, ms_0 = ev(net), ev(net0)
        n_launch = 2 * fl["evals"]
        per_launch_ms = (ms_T - ms_0) / n_launch
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if w["dtype"] == "bf16":
            # The dominant kernel is the whole-solve launch of tc_staged_gemm: ALL T Euler steps (ERR + GRAD tiles of every
            # layer) in one launch.  Timed live: a T-step jpcb_pc_infer minus a 0-step one (operand preparation only), CUDA
            # events on the launching stream.  From arbitrary activities the schedule is dense (no wavefront shortcut), so
            # the launch's algorithmic FLOPs are T x per_eval.  The timed window is ~0.1 s at 1.9+ GHz, not the seconds-long
            # power-capped regime: the denominator is the measured BURST cuBLAS figure; the sustained one is given beside it.
            burst = peaks.get("bf16_tflops", 1590.0)
            sust = peaks.get("bf16_tflops_sustained", 1400.0)
            src = "measured burst cuBLAS bf16 (MEASURED_PEAKS.json)" if peaks else "fallback 1.59 PF burst (B200_PROFILING.md)"
            launch_ms = ms_T - ms_0
            fl_launch = fl["evals"] * fl["per_eval"]
            ach = fl_launch / (launch_ms * 1e-3) / 1e12
            traffic = None
            try:  # dram__bytes_read.sum + dram__bytes_write.sum of this launch, from the committed ncu --set full capture
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]["bytes_per_launch"]
            except Exception:
                pass
            roof = {"bound": "tensor", "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst, "traffic": traffic,
                    "traffic_source": "ncu --set full capture of the same launch, committed under profiles/ (not a live counter)",
                    "kernel": "tc_staged_gemm whole-solve launch (all T Euler steps: ERR + GRAD tiles of every layer, one "
                              "dependency-tracked persistent stream)",
                    "launch_ms": launch_ms, "flops_per_launch": fl_launch, "peak_source": src,
                    "peak_sustained": sust, "frac_of_sustained": ach / sust,
                    "per_phase_ms": per_launch_ms,
                    "whole_step_tflops_executed": fl["executed_total"] / (ms / args.steps * 1e-3) / 1e12,
                    "whole_step_frac_of_burst": fl["executed_total"] / (ms / args.steps * 1e-3) / 1e12 / burst}
        else:
            # narrow fp32 configs: latency-bound; report algorithmic HBM bytes of one phase launch
            dims = [w["d_in"]] + [w["width"]] * (w["depth"] - 1) + [w["d_out"]]
            wbytes = 4 * sum(dims[l] * dims[l + 1] for l in range(1, w["depth"]))
            abytes = 4 * w["B"] * sum(dims[1:]) * 2
            peak = peaks.get("hbm_gbs", 6650.0)
            ach = (wbytes + abytes) / (per_launch_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "kernel": "warp_grouped_gemm (narrow-layer fp32 path; one inference phase = all layers' GEMMs + fused epilogue)",
                    "launch_ms": per_launch_ms,
                    "bytes_per_launch": wbytes + abytes,
                    "whole_step_tflops": fl["total"] / (ms / args.steps * 1e-3) / 1e12}

    # ---- CPU baseline (rank 0, N == 1 only): the oracle timed on the host cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows = min(w["B"], args.cpu_rows)
        cstep = cpu_reference_step_fn(w, rows)
        cstep()
        reps, t0 = 0, time.perf_counter()
        while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 50):
            cstep()
            reps += 1
        cdt = time.perf_counter() - t0
        cpu = {"value": rows * reps / cdt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{reps} steps of {rows} of {w['B']} rows, fp32 torch-CPU autograd restatement of jpc.make_pc_step"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None,
            "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": w["desc"], "global_batch": B_global, "inference_steps": T, "solver": w["solver"],
                       "parallelism": f"dp{world}", "l2": "working set (activities + operand copies > 126 MB L2) exceeds L2; no flush"
                       if args.workload == "c4" else "working set fits L2 (latency-bound config); no flush",
                       "algorithmic_gflop_per_step": fl["total"] / 1e9, "executed_gflop_per_step": fl["executed_total"] / 1e9,
                       "wavefront": "from the feed-forward init only layers the error wavefront has reached are evaluated "
                                    "(same result; executed < dense FLOPs)"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": x_pin.numel() * 4 + y_pin.numel() * 4,
                    "d2h_bytes_per_step": 4,
                    "d2h": "non-blocking copy of the step's loss into pinned host memory every step, drained at the end of the "
                           "timed region", "warmup": e2e_first.get("warmup", E2E_WARM),
                    "plan_captures_in_timed_region": e2e_first.get("plan_captures_in_timed_region")},
            "e2e_blocking_read": {"value": e2e_blocking_val, "unit": UNIT,
                                  "what": "the same loop with the host blocking on every step's loss (float(loss))"},
            "e2e_graph": e2e_graph,
            "gpu_launches": int(launches), "clocks": clocks, "launch_plans": J.plan_cache_stats(), "roofline": roof,
            "cpu_baseline": cpu,
            "loss": loss_val,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # (leave without tearing the communicators down: destroying a process group whose collective was captured into a
        #  still-alive CUDA graph -- PCTrainer under data parallelism -- blocked at exit on this stack; nothing is left to flush)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
