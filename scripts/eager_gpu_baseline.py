#!/usr/bin/env python
"""Second comparator of SURVEY.md §8d: the reference's forward as PyTorch-eager library kernels ON the B200
(cuBLAS / ATen; the reference ships no GPU kernels of its own), CUDA-event timed.  Uses the torch restatement of the
reference forward (oracle/torch_port.py — test infrastructure; the reference itself does not travel to the GPU box),
hm_0 architecture, fp32 / TF32-allowed / bf16-autocast, at the reference's own batch size (256) and at 65 536.

    python scripts/eager_gpu_baseline.py > gpurun_out/eager_gpu.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openmpl_b200 import spec, synth
from oracle import torch_port

kw = dict(num_joints=17, embed_dim_ratio=32, num_heads=8, depth=12, num_views=4, drop_path_rate=0.1, **spec.HM0_FLAGS)
cfg = spec.make_config(**kw)
dev = torch.device("cuda", 0)
weights = {k: torch.from_numpy(v).to(dev) for k, v in synth.named_weights(spec.param_spec(cfg), seed=0).items()}
rig = synth.make_rig(4)
for B in [int(a) for a in sys.argv[1:]] or (256, 16384, 65536):
    batch = synth.make_batch(B, rig, seed=1)
    x = [torch.from_numpy(batch[k]).to(dev) for k in ("poses", "rays", "centers")]
    for mode in ("fp32", "tf32", "bf16-autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = mode == "tf32"

        def step():
            if mode == "bf16-autocast":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return torch_port.forward(weights, cfg, *x)
            return torch_port.forward(weights, cfg, *x)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        n = 5 if B > 1000 else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps({"comparator": "PyTorch-eager restatement of the reference forward on B200", "mode": mode, "batch": B,
                          "ms_per_forward": ms, "poses_per_s": B / (ms / 1000.0)}), flush=True)
