"""Small forwards for `compute-sanitizer --tool racecheck`: bf16 with the fused QKV + attention kernel and the residual-emit
epilogues, bf16 with the stand-alone attention (uncertainty weights), and the split-operand fp32-grade mode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openmpl_b200 import spec, synth
from openmpl_b200.models.multiview_mpl_b200 import MultiView_MPL

for extra, prec in ((dict(), "bf16"), (dict(confidence_as_attention_uncertainty_weight=True), "bf16"), (dict(), "tf32")):
    kw = dict(num_joints=17, embed_dim_ratio=32, num_heads=8, drop_path_rate=0.1, depth=2, num_views=4, **extra, **spec.HM0_FLAGS)
    cfg = spec.make_config(**kw)
    w = synth.named_weights(spec.param_spec(cfg), seed=0)
    batch = synth.make_batch(40, synth.make_rig(cfg.V), seed=2)
    m = MultiView_MPL(**kw, precision=prec, graph_batch=0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    m = m.cuda().eval()
    a = [torch.from_numpy(batch[k]).cuda() for k in ("poses", "rays", "centers")]
    with torch.no_grad():
        out = m(a[0], rays=a[1], centers=a[2])
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    print("sanitize_race:", prec, extra, "ok", flush=True)
print("sanitize_race: done")
