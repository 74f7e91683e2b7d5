"""Small forwards of every precision / several flag sets, meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openmpl_b200 import evaluate, spec, synth
from openmpl_b200.models.multiview_mpl_b200 import MultiView_MPL

CASES = [
    dict(depth=2, num_views=4, **spec.HM0_FLAGS),
    dict(depth=2, num_views=3, **spec.CHOSEN_FLAGS),
    dict(depth=1, num_views=5, **spec.HM0_FLAGS),        # pose-aligned tiling of the fused QKV + attention kernel, permuted planes
    dict(depth=1, num_views=7, **spec.CHOSEN_FLAGS),     # 68-wide head pairs, 28 rows per lane quarter
    dict(depth=1, num_views=5, confidence_as_attention_uncertainty_weight=True, confidence_in_FPT=True, input_rays_as_token=True),
    dict(depth=2, num_views=2, FPT_blocks_view_keypoint_tokens=True, pose_3d_emb_learnable=True),
]
for kw in CASES:
    kw = dict(num_joints=17, embed_dim_ratio=32, num_heads=8, drop_path_rate=0.1, **kw)
    cfg = spec.make_config(**kw)
    w = synth.named_weights(spec.param_spec(cfg), seed=0)
    for B in (1, 37, 300):
        batch = synth.make_batch(B, synth.make_rig(cfg.V), seed=2)
        for prec in ("fp32", "tf32", "bf16"):
            # batches up to graph_batch replay a captured CUDA graph; B = 300 runs kernel by kernel
            m = MultiView_MPL(**kw, precision=prec, graph_batch=64)
            m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
            m = m.cuda().eval()
            a = [torch.from_numpy(batch[k]).cuda() for k in ("poses", "rays", "centers")]
            with torch.no_grad():
                out = m(a[0], rays=a[1], centers=a[2])
            torch.cuda.synchronize()
            assert torch.isfinite(out).all()
            if B == 300 and prec == "bf16":                  # pipelined host staging: pinned host inputs over several chunks
                m.set_chunk_poses(128)
                m.pipeline_first_poses = 50
                h = [torch.from_numpy(batch[k]).pin_memory() for k in ("poses", "rays", "centers")]
                with torch.no_grad():
                    out2 = m(h[0], rays=h[1], centers=h[2])
                torch.cuda.synchronize()
                assert torch.equal(out, out2)
print("eval loop")
evaluate.run(arch="cmu0", views=2, poses=700, micro_batch=256, precision="bf16")
print("sanitize_small: done")
