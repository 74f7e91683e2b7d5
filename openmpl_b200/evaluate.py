"""Large-batch evaluation loop on synthetic MHP-style data — the device-resident counterpart of the reference's
`validate()` (`MPL/lib/core/function_mpl.py:293-649`) for BASELINE.json configs 3-5.

    python -m openmpl_b200.evaluate --arch hm0 --views 4 --poses 10000000               # config 5 on one GPU
    torchrun --nproc-per-node 8 -m openmpl_b200.evaluate --arch cmu0 --views 5 ...      # config 3: ranks shard the poses
    python -m openmpl_b200.evaluate --arch kptok --views 8 --depth 12 --poses 262144    # config 4: view sweep

Per micro-batch, entirely on the device: `mpl_synth_project` (3D poses from the counter-based generator keyed by the
GLOBAL pose index, projected through the V calibrations) -> `mpl_build_inputs` (clip, confidence zeroing, screen
normalisation, rays) -> `mpl_forward` -> `mpl_mpjpe_accumulate` + `mpl_pmpjpe_accumulate` (fp64 running sums).  Predictions never leave the GPU;
ranks own contiguous slices of the index range, and one all-reduce of the 11 J + 1 (and J + 3 Procrustes) sums at the end
gives the MPJPE the reference's `evaluate()` would log (`function_mpl.py:670-687`) and the P-MPJPE of
`pose_utils.py:61-143`.  Prints one JSON line from rank 0.
"""
from __future__ import annotations

import argparse
import json
import sys

import torch

from . import dist as mdist, inputs, metric, spec, synth
from .models.multiview_mpl_b200 import MultiView_MPL

ARCHS = {
    # constructor flag sets of the shipped YAMLs / ablations
    "hm0": dict(spec.HM0_FLAGS), "cmu0": dict(spec.HM0_FLAGS), "chosen": dict(spec.CHOSEN_FLAGS),
    "kptok": dict(pose_3d_emb_learnable=True, confidence_input_as_third=True, FPT_blocks_view_keypoint_tokens=True),
}
DEFAULT_DEPTH = {"hm0": 12, "chosen": 12, "kptok": 12, "cmu0": 2}
DEFAULT_RIG = {"hm0": "h36m", "chosen": "h36m", "kptok": "h36m", "cmu0": "cmu"}


def run(arch="hm0", views=4, depth=None, poses=1 << 20, micro_batch=65536, precision="bf16", rig=None, seed=1,
        weight_seed=0, warmup=1):
    rank, world, local = mdist.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    depth = DEFAULT_DEPTH[arch] if depth is None else depth
    kw = dict(num_joints=17, embed_dim_ratio=32, num_heads=8, depth=depth, num_views=views, drop_path_rate=0.1, **ARCHS[arch])
    cfg = spec.make_config(**kw)
    model = MultiView_MPL(**kw, precision=precision)
    weights = synth.named_weights(spec.param_spec(cfg), seed=weight_seed)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model = model.to(dev).eval()
    mdist.broadcast_state(model)                       # replicas bit-identical (they already are: same seed)
    the_rig = synth.make_rig(views, rig or DEFAULT_RIG[arch])
    start, stop = mdist.shard_range(poses, rank, world)
    acc = metric.MpjpeAccumulator(cfg.J, output_in_meter=True, device=dev)
    pacc = metric.PmpjpeAccumulator(cfg.J, output_in_meter=True, device=dev)

    def step(s0, n):
        pix, target, calib = inputs.synth_project(n, the_rig, seed=seed, start=s0, device=dev)
        p, r, c = inputs.build_inputs(pix, calib)
        with torch.no_grad():
            out = model(p, rays=r, centers=c)
        out = out[0] if isinstance(out, tuple) else out
        acc.update(out, target)
        pacc.update(out, target)

    for _ in range(warmup):                            # allocate the workspace, pack the weights
        step(start, min(micro_batch, max(stop - start, 1)))
    acc.acc.zero_()
    pacc.acc.zero_()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s0 in range(start, stop, micro_batch):
        step(s0, min(micro_batch, stop - s0))
    acc.all_reduce()                                   # the only collectives of the whole job:
    pacc.all_reduce()                                  # 11 J + 1 and J + 3 doubles
    e1.record()
    torch.cuda.synchronize()
    ms = mdist.max_over_ranks(e0.elapsed_time(e1), dev)
    res, pres = acc.result(), pacc.result()
    line = None
    if rank == 0:
        line = {
            "metric": "poses/sec end-to-end synthetic evaluation (generate + project + build inputs + forward + MPJPE)",
            "value": poses / (ms / 1000.0), "unit": "poses/s", "n_gpus": world, "ms_total": ms, "poses": res["n"],
            "dtype": precision, "data": "synthetic (device generator keyed by global pose index)",
            "config": {"workload": f"{arch} V={views} depth={depth} D={cfg.fpt_dim} tokens={cfg.fpt_tokens}",
                       "micro_batch": micro_batch, "rig": rig or DEFAULT_RIG[arch], "flops_per_pose": spec.flops_per_pose(cfg)},
            "mpjpe_cm": {"absolute": res["mpjpe_abs"], "root_relative": res["mpjpe_rel"], "procrustes_aligned": pres["p_mpjpe"],
                         "note": "random-init weights: exercises the metric path, not a trained accuracy"},
            "tflops": poses / (ms / 1000.0) * spec.flops_per_pose(cfg) / 1e12 / world,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return line


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--arch", default="hm0", choices=sorted(ARCHS))
    p.add_argument("--views", type=int, default=4)
    p.add_argument("--depth", type=int, default=None)
    p.add_argument("--poses", type=int, default=1 << 20, help="global number of poses (sharded over ranks)")
    p.add_argument("--micro-batch", type=int, default=65536)
    p.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    p.add_argument("--rig", default=None, choices=[None, "h36m", "cmu"])
    p.add_argument("--seed", type=int, default=1)
    a = p.parse_args(argv)
    run(a.arch, a.views, a.depth, a.poses, a.micro_batch, a.precision, a.rig, a.seed)


if __name__ == "__main__":
    sys.exit(main())
