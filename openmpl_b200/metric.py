"""K6 host side: MPJPE accumulators on the device + their reduction over ranks.

Mirrors `evaluate()` of the reference (`MPL/lib/core/function_mpl.py:670-687`) and `calc_mpjpe` /
`calc_distance_per_dim` (`MPL/lib/core/evaluate.py:91-125`): unit rule (x100 when OUTPUT_IN_METER), root-relative
variant, `joints_3d_conf <= 0` masking with NaN semantics.  Predictions never leave the GPU: each batch adds into
`11 J + 1` fp64 running sums (layout in include/mpl_b200.h) and ranks are combined with ONE all-reduce at the end.
`PmpjpeAccumulator` does the same for the Procrustes-aligned error (`MPL/lib/utils/pose_utils.py:61-143`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def acc_len(J: int) -> int:
    return 11 * J + 1


class MpjpeAccumulator:
    def __init__(self, num_joints: int = 17, output_in_meter: bool = True, device=None):
        self.J = num_joints
        self.unit = 100.0 if output_in_meter else 1.0
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.acc = torch.zeros(acc_len(num_joints), dtype=torch.float64, device=self.device)

    def update(self, pred: torch.Tensor, gt: torch.Tensor, conf3d: torch.Tensor | None = None, room: dict | None = None):
        """pred, gt: [B, J, 3] fp32 on the device; conf3d: None, [B, J], [B, J, 1] or [B, J, 3] (the dataset's
        `joints_3d_conf`, `function_mpl.py:489-490,682-684`).  room: the `meta` entries of a room-normalised dataset --
        {'room_x_scale', 'room_center'} when 'room_scaled_equal', else {'room_x_scale', 'room_y_scale'} -- for the
        un-scaling `validate()` applies before it stores predictions and targets (`function_mpl.py:476-488`)."""
        B = pred.shape[0]
        affine = None
        if room is not None:
            if "room_center" in room:                     # 'room_scaled_equal': v * s + centre on every axis
                s = float(room["room_x_scale"])
                c = [float(x) for x in room["room_center"]]
                vals = [s, s, s] + c
            else:                                         # x and y scaled separately, z untouched
                vals = [float(room["room_x_scale"]), float(room["room_y_scale"]), 1.0, 0.0, 0.0, 0.0]
            import ctypes
            affine = (ctypes.c_float * 6)(*vals)
        pred = pred.to(self.device, torch.float32).contiguous()
        gt = gt.to(self.device, torch.float32).contiguous()
        if conf3d is not None:
            c = conf3d.to(self.device, torch.float32)
            if c.dim() == 2:
                c = c[:, :, None]
            conf3d = c.expand(B, self.J, 3).contiguous()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().mpl_mpjpe_accumulate(pred.data_ptr(), gt.data_ptr(),
                                                       conf3d.data_ptr() if conf3d is not None else None, B, self.J,
                                                       self.unit, affine, self.acc.data_ptr(), stream))

    def all_reduce(self):
        """Sum the accumulators over all ranks (NCCL over NVLink on the GPU box; one call of `11 J + 1` doubles)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
        return self

    def result(self) -> dict:
        return finalize(self.acc.detach().cpu().numpy(), self.J)


class PmpjpeAccumulator:
    """Procrustes-aligned MPJPE (P-MPJPE): every prediction is first aligned to its ground truth with
    `PoseUtils.procrustes` of the reference (`MPL/lib/utils/pose_utils.py:61-143`; similarity transform, reflection
    'best' by default), then scored like `calc_mpjpe`.  `J + 3` fp64 running sums on the device, one all-reduce."""

    REFLECTION = {"best": -1, False: 0, True: 1}

    def __init__(self, num_joints: int = 17, output_in_meter: bool = True, scaling: bool = True, reflection="best",
                 device=None):
        self.J = num_joints
        self.unit = 100.0 if output_in_meter else 1.0
        self.scaling = bool(scaling)
        self.reflection = self.REFLECTION[reflection]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.acc = torch.zeros(num_joints + 3, dtype=torch.float64, device=self.device)

    def update(self, pred: torch.Tensor, gt: torch.Tensor):
        """pred, gt: [B, J, 3] fp32 on the device."""
        B = pred.shape[0]
        pred = pred.to(self.device, torch.float32).contiguous()
        gt = gt.to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().mpl_pmpjpe_accumulate(pred.data_ptr(), gt.data_ptr(), B, self.J, self.unit,
                                                        int(self.scaling), self.reflection, self.acc.data_ptr(), stream))

    def all_reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
        return self

    def result(self) -> dict:
        return finalize_pmpjpe(self.acc.detach().cpu().numpy(), self.J)


def finalize_pmpjpe(acc: np.ndarray, J: int) -> dict:
    """Running sums -> per-joint P-MPJPE, its mean, mean normalised residual d and mean scale."""
    acc = np.asarray(acc, dtype=np.float64)
    n = acc[J + 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        out = {"n": int(round(n)), "pjpe_aligned": acc[0:J] / n, "residual": float(acc[J] / n), "scale": float(acc[J + 1] / n)}
    out["p_mpjpe"] = float(out["pjpe_aligned"].mean())
    return out


def finalize(acc: np.ndarray, J: int) -> dict:
    """Running sums -> the quantities `evaluate()` logs: per-joint MPJPE, its mean, per-dim distances."""
    acc = np.asarray(acc, dtype=np.float64)
    n = acc[11 * J]
    cnt = acc[8 * J:11 * J].reshape(J, 3)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = {
            "n": int(round(n)),
            "pjpe_abs": acc[0:J] / n,
            "pjpe_rel": acc[J:2 * J] / n,
            "dist_abs": acc[2 * J:5 * J].reshape(J, 3) / cnt,      # nanmean over unmasked entries
            "dist_rel": acc[5 * J:8 * J].reshape(J, 3) / cnt,
        }
    out["mpjpe_abs"] = float(out["pjpe_abs"].mean())
    out["mpjpe_rel"] = float(out["pjpe_rel"].mean())
    return out
