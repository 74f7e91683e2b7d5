#!/usr/bin/env python
"""bench.py — poses/s of the MPL lifter forward (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|tf32|fp32]

One "step" = one forward of the H36M 4-view 17-joint `hm_0` lifter (depth 12, D = 1088, 114 M parameters) over a
batch of 65 536 synthetic poses per GPU (BASELINE.json configs[1]).  Ranks shard the pose index range; there is no
data-path collective, only one all-reduce of the MPJPE accumulators after the timed region.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "poses/sec MPL forward (H36M 4-view, 17 joints)"
ARCH = dict(num_joints=17, embed_dim_ratio=32, num_heads=8, depth=12, num_views=4, drop_path_rate=0.1)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--precision", default=os.environ.get("MPL_BENCH_PRECISION", "bf16"), choices=["bf16", "tf32", "fp32"])
    p.add_argument("--batch", type=int, default=65536, help="poses per GPU per step")
    p.add_argument("--cpu-batch", type=int, default=1024, help="poses per CPU-baseline forward")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--depth", type=int, default=12)
    return p.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def workload(args):
    from openmpl_b200 import spec
    kw = dict(ARCH, depth=args.depth, **spec.HM0_FLAGS)
    return kw, spec.make_config(**kw)


def cpu_forward_timer(args, steps, warmup):
    """The reference forward's CPU path timed on the host cores: the torch-CPU restatement (same ATen/MKL calls the
    reference module makes, fp32, all cores) — the reference itself is Python and does not travel to the GPU box."""
    import torch
    from openmpl_b200 import spec, synth
    from oracle import torch_port                                   # cpu_baseline / --impl reference legs only
    torch.set_num_threads(os.cpu_count() or 1)
    kw, cfg = workload(args)
    weights = synth.named_weights(spec.param_spec(cfg), seed=0)
    batch = synth.make_batch(args.cpu_batch, synth.make_rig(cfg.V), seed=1)
    p = {k: torch.from_numpy(v) for k, v in weights.items()}
    x = [torch.from_numpy(batch[k]) for k in ("poses", "rays", "centers")]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        torch_port.forward(p, cfg, *x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def run_reference(args):
    """--impl reference: the reference algorithm's CPU path (oracle port; the reference itself is pure Python/PyTorch and
    does not travel to the GPU box) on all host cores; each step = one forward over a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warmup = max(1, min(args.steps, 100)), max(1, min(args.warmup, 5))   # K timed forwards, W untimed
    times = cpu_forward_timer(args, steps, warmup)
    total = sum(times)
    value = args.cpu_batch * len(times) / total
    kw, cfg = workload(args)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "poses/s", "n_gpus": args.gpus, "steps": len(times),
        "warmup": warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"H36M 4-view 17-joint hm_0 lifter forward (MultiSPT, Conf3rd, Raytoken, Add3dEncRays), depth {cfg.depth}, "
                               f"D={cfg.fpt_dim}, bounded sample of {args.cpu_batch} poses per step",
                   "batch_per_step": args.cpu_batch, "views": cfg.V, "joints": cfg.J, "note": "torch-CPU fp32 restatement of MultiView_MPL.forward (same ATen calls as the reference) on all host cores"},
        "cpu_baseline": {"value": value, "unit": "poses/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} forwards of {args.cpu_batch} poses"},
        "e2e": {"value": value, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().strip().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the samples taken under load (upper half of the power readings)
        order = np.argsort(pw)[len(pw) // 2:]
        return {"sm_mhz": float(np.median(np.asarray(sm)[order])), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from openmpl_b200 import dist as mdist, metric, spec, synth
    from openmpl_b200.models.multiview_mpl_b200 import MultiView_MPL

    rank, world, local = mdist.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    kw, cfg = workload(args)
    B = args.batch
    pk, pk_src = peaks()

    # ---- model: random-init weights of the named architecture (seeded, identical on every rank) ----
    weights = synth.named_weights(spec.param_spec(cfg), seed=0)
    impl = {}
    if os.environ.get("MPL_GEMM_CTA_GROUP"):
        impl["gemm_cta_group"] = int(os.environ["MPL_GEMM_CTA_GROUP"])
    if os.environ.get("MPL_LN_FUSION"):
        impl["ln_fusion"] = bool(int(os.environ["MPL_LN_FUSION"]))
    model = MultiView_MPL(**kw, precision=args.precision, **impl)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model = model.to(dev).eval()
    if os.environ.get("MPL_CHUNK"):
        model.set_chunk_poses(int(os.environ["MPL_CHUNK"]))

    # ---- synthetic inputs of this rank's shard of the global pose range ----
    start, _ = mdist.shard_range(B * world, rank, world)
    rig = synth.make_rig(cfg.V, "h36m")
    batch = synth.make_batch(B, rig, seed=1, start=start)
    host = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("poses", "rays", "centers", "target")}
    devin = {k: v.to(dev) for k, v in host.items()}
    h2d = world * sum(host[k].numel() * 4 for k in ("poses", "rays", "centers"))     # whole job, all ranks
    d2h = world * B * cfg.J * 3 * 4

    def step_device():
        with torch.no_grad():
            return model(devin["poses"], rays=devin["rays"], centers=devin["centers"])

    host_out = torch.empty((B, cfg.J, 3), dtype=torch.float32).pin_memory()

    def step_e2e():
        with torch.no_grad():
            out = model(host["poses"], rays=host["rays"], centers=host["centers"])     # H2D inside the module call
        host_out.copy_(out, non_blocking=True)                                          # D2H of the step's result (pinned)
        torch.cuda.synchronize()
        return host_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity spot-check against the oracle (checker only; outside every timed region) ----
    out = step_device()
    parity = None
    if rank == 0:
        from oracle import mpl_oracle
        idx = np.arange(0, B, max(1, B // 8))[:8]
        ref = mpl_oracle.forward(weights, cfg, batch["poses"][idx], batch["rays"][idx], batch["centers"][idx])
        got = out[torch.from_numpy(idx).to(dev)].cpu().numpy()
        scale = float(np.abs(ref).max())
        tgt = batch["target"][idx].astype(np.float64)
        mp = lambda p: float(np.sqrt(((p - tgt) ** 2).sum(-1)).mean()) * 1000.0
        parity = {"max_abs_err_over_scale": float(np.abs(got - ref).max()) / scale, "poses_checked": int(len(idx)),
                  "delta_mpjpe_mm": abs(mp(got.astype(np.float64)) - mp(ref))}

    # ---- timed region 1: device-resident inputs (value) ----
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        step_device()
        launches += model.last_launches
    e1.record()
    barrier()
    ms = mdist.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if sampler else None
    value = world * args.steps * B / (ms / 1000.0)

    # ---- timed region 2: end to end through the module call with host buffers (e2e) ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    ms_e2e = mdist.max_over_ranks((time.perf_counter() - t0) * 1000.0, dev)
    barrier()
    e2e_value = world * args.steps * B / (ms_e2e / 1000.0)

    # ---- per-kernel breakdown with CUDA events on the launch stream (roofline) ----
    model.set_profile(True)
    prof_steps = 2
    agg = {}
    for _ in range(prof_steps):
        step_device()
        for k, (t, n) in model.profile().items():
            a = agg.setdefault(k, [0.0, 0])
            a[0] += t; a[1] += n
    model.set_profile(False)
    gemm_cats = ["fpt_gemm_qkv", "fpt_gemm_proj", "fpt_gemm_fc1", "fpt_gemm_fc2"]
    D, Hf, M = cfg.fpt_dim, cfg.fpt_hidden, B * cfg.fpt_tokens
    flops_per_launch = {"fpt_gemm_qkv": 2.0 * M * 3 * D * D, "fpt_gemm_proj": 2.0 * M * D * D,
                        "fpt_gemm_fc1": 2.0 * M * Hf * D, "fpt_gemm_fc2": 2.0 * M * D * Hf}
    chunk_poses = int(os.environ.get("MPL_CHUNK", "0")) or int(model.chunk_poses())
    chunks = -(-B // chunk_poses)
    roofline, breakdown = None, {}
    try:      # per-launch DRAM traffic of the GEMM launches from the committed `ncu --set full` capture (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        traffic = {}
    tot_ms = sum(a[0] for a in agg.values())
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        breakdown[k] = {"ms_per_step": t / prof_steps, "launches_per_step": n // prof_steps, "share": t / tot_ms if tot_ms else 0}
    if all(c in agg for c in gemm_cats):
        g_ms = sum(agg[c][0] for c in gemm_cats)
        g_n = sum(agg[c][1] for c in gemm_cats)
        # algorithmic FLOPs of the launches timed: each category launches (depth+1) * chunks times per step on M / chunks rows
        g_flops = sum(flops_per_launch[c] / chunks * agg[c][1] for c in gemm_cats)
        achieved = g_flops / (g_ms / 1000.0) / 1e12
        # tf32-named mode = split bf16 operands: three bf16 MMAs per algorithmic product
        peak = pk.get("bf16_tflops_sustained", 1400.0) / (3.0 if args.precision == "tf32" else 1.0)
        roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (QKV/proj/fc1/fc2 of the FPT)", "achieved": achieved,
                    "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic.get("gemm_tcgen05_kernel", {}).get("bytes_per_launch"),
                    "traffic_note": traffic.get("gemm_tcgen05_kernel", {}).get("note"),
                    "algorithmic_flops_per_launch": g_flops / g_n,
                    "peak_source": f"{pk_src} bf16_tflops_sustained" + (" / 3 (split bf16 hi/lo operands)" if args.precision == "tf32" else ""),
                    "avg_launch_ms": g_ms / g_n, "launches_timed": g_n, "share_of_step": g_ms / tot_ms}
    elif agg:
        k = max(agg, key=lambda c: agg[c][0])
        roofline = {"bound": "tensor", "kernel": k, "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None,
                    "traffic": None, "note": "fp32 CUDA-core path: no tensor-core kernel in this mode"}

    # ---- HBM-bound kernels: algorithmic bytes per launch / measured launch time vs the measured copy bandwidth ----
    Bc = min(B, chunk_poses)
    rows_f = Bc * cfg.fpt_tokens
    esz = 2 if args.precision == "bf16" else 4
    alg_bytes = {
        "embed": Bc * cfg.V * (cfg.J * 12 + cfg.J * cfg.d * 4),
        "token_build": Bc * cfg.V * (cfg.J * cfg.d * 4 + 2 * cfg.J * 12 + 12 + cfg.tok_w * 4),
        "fpt_attention": rows_f * (3 * D + D) * esz,
        "fpt_layernorm": rows_f * (D * 4 + D * esz),
        "head": Bc * (cfg.V * cfg.E * 4 + cfg.J * 12),
    }
    memory_kernels = {}
    for k, nbytes in alg_bytes.items():
        if k in agg and agg[k][1]:
            ms_l = agg[k][0] / agg[k][1]
            gbs = nbytes / (ms_l * 1e-3) / 1e9
            memory_kernels[k] = {"avg_launch_ms": ms_l, "algorithmic_bytes_per_launch": nbytes, "achieved_gbs": gbs,
                                 "frac_of_hbm_peak": gbs / pk.get("hbm_gbs", 6650.0)}

    # ---- the collectives: MPJPE / P-MPJPE accumulators all-reduced over ranks (NCCL) ----
    acc = metric.MpjpeAccumulator(cfg.J, output_in_meter=True, device=dev)
    acc.update(out, devin["target"])
    acc.all_reduce()
    res = acc.result()
    pacc = metric.PmpjpeAccumulator(cfg.J, output_in_meter=True, device=dev)
    pacc.update(out, devin["target"])
    pacc.all_reduce()
    pres = pacc.result()

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times = cpu_forward_timer(args, steps=3, warmup=1)
        v = args.cpu_batch * len(times) / sum(times)
        cpu = {"value": v, "unit": "poses/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{len(times)} forwards of {args.cpu_batch} poses (torch-CPU fp32 restatement of the reference forward, all cores)"}

    if rank == 0:
        flops = spec.flops_per_pose(cfg)
        line = {
            "metric": METRIC, "value": value, "unit": "poses/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"H36M 4-view 17-joint hm_0 lifter forward (MultiSPT, Conf3rd, Raytoken, Add3dEncRays), depth {cfg.depth}, "
                                   f"D={cfg.fpt_dim}, batch {B} per GPU", "batch_per_gpu": B, "views": cfg.V, "joints": cfg.J,
                       "parallelism": f"pose-sharded x{world}, no data-path collective",
                       "l2": "activation working set per step (GBs) far exceeds the 126 MB L2; no explicit flush",
                       "flops_per_pose": flops, "gemm_cta_group": int(os.environ.get("MPL_GEMM_CTA_GROUP", "0")) or None},
            "e2e": {"value": e2e_value, "unit": "poses/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": roofline, "whole_path_tflops": value / world * flops / 1e12,
            "breakdown": breakdown, "memory_bound_kernels": memory_kernels, "hbm_peak_gbs": pk.get("hbm_gbs"),
            "cpu_baseline": cpu, "parity": parity,
            "mpjpe_cm": {"absolute": res["mpjpe_abs"], "root_relative": res["mpjpe_rel"], "procrustes_aligned": pres["p_mpjpe"],
                         "poses": res["n"],
                         "note": "random-init weights: the values only exercise the accumulators + all-reduce"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
