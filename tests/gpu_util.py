"""Helpers shared by the -m gpu parity tests (the oracle is only ever the checker)."""
import numpy as np
import torch

from oracle import mpl_oracle
from openmpl_b200.models import multiview_mpl_b200 as mb

# Stated tolerances (per-coordinate max abs error / output scale), BASELINE.json north_star:
#   fp32  : the path that carries the "<= 1e-3 of scale, dMPJPE <= 0.1 mm" claim (measured ~1e-6)
#   tf32  : single-pass tcgen05 kind::tf32 projections; measured 0.6e-3 .. 2.1e-3 of scale on the parity cases, i.e. it
#           does NOT always meet 1e-3 (SURVEY.md §7-H5 predicted this) -> stated bound 4e-3, opt-in only
#   bf16  : bf16 tensor-core operands, fp32 accumulate / LayerNorm / softmax / residual -> stated bound 3e-2
TOL = {"fp32": 2e-5, "tf32": 4e-3, "bf16": 3e-2}
DMPJPE_MM = {"fp32": 0.1, "tf32": 0.5, "bf16": 1.5}     # |MPJPE(new) - MPJPE(reference)| in mm (targets in metres)


def build_module(kw, weights, precision, device="cuda"):
    m = mb.MultiView_MPL(**kw, precision=precision)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}, strict=True)
    return m.to(device).eval()


def run_module(m, batch, packed=False, device="cuda"):
    V = batch["poses"].shape[1]
    if packed:
        args = [torch.from_numpy(batch[k]).to(device) for k in ("poses", "rays", "centers")]
    else:
        args = [[torch.from_numpy(np.ascontiguousarray(batch[k][:, v])).to(device) for v in range(V)]
                for k in ("poses", "rays", "centers")]
    with torch.no_grad():
        out = m(args[0], rays=args[1], centers=args[2])
    torch.cuda.synchronize()
    if isinstance(out, tuple):
        return [out[0].cpu().numpy()] + [o.cpu().numpy() for o in out[1]]
    return [out.cpu().numpy()]


def oracle_outputs(cfg, weights, batch):
    out = mpl_oracle.forward(weights, cfg, batch["poses"], batch["rays"], batch["centers"])
    return [out[0]] + list(out[1]) if isinstance(out, tuple) else [out]


def rel_err(a, ref):
    scale = max(float(np.abs(ref).max()), 1e-6)
    return float(np.abs(a.astype(np.float64) - ref).max()) / scale


def mpjpe_mm(pred, target):
    return float(np.sqrt(((pred.astype(np.float64) - target) ** 2).sum(-1)).mean()) * 1000.0
