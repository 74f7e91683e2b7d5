#!/usr/bin/env python
"""Per-kernel-category time of one forward for any architecture:  python scripts/profile_arch.py kptok 4 [batch] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openmpl_b200 import evaluate, spec, synth
from openmpl_b200.models.multiview_mpl_b200 import MultiView_MPL

arch, V = sys.argv[1], int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
kw = dict(num_joints=17, embed_dim_ratio=32, num_heads=8, depth=evaluate.DEFAULT_DEPTH[arch], num_views=V, drop_path_rate=0.1,
          **evaluate.ARCHS[arch])
cfg = spec.make_config(**kw)
m = MultiView_MPL(**kw, precision=prec)
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.named_weights(spec.param_spec(cfg), seed=0).items()})
m = m.cuda().eval()
batch = synth.make_batch(B, synth.make_rig(V, evaluate.DEFAULT_RIG[arch]), seed=1)
x = [torch.from_numpy(batch[k]).cuda() for k in ("poses", "rays", "centers")]
with torch.no_grad():
    for _ in range(3):
        m(x[0], rays=x[1], centers=x[2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m(x[0], rays=x[1], centers=x[2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    m.set_profile(True)
    m(x[0], rays=x[1], centers=x[2])
    prof = m.profile()
print(f"{arch} V={V} B={B} {prec}: {ms:.2f} ms/forward = {B / ms * 1e3:,.0f} poses/s, {m.last_launches} launches")
for k, (t, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"   {k:16s} {t:8.3f} ms  {n:4d} launches")
