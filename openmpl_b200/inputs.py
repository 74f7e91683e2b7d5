"""N2: batched input builder — raw detector output + calibrations -> the model's (poses, rays, centers).

Replaces the per-sample numpy of `JointsDataset_MPL.__getitem__`
(`MPL/lib/dataset/joints_dataset_mpl.py:615-648,701-715,762-772,817-820,872-904`) with one streaming kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def pack_calibration(R, t, f, c, image_size) -> np.ndarray:
    """[V, 18] float64 rows: R (9, row-major world->cam), t (3), fx, fy, cx, cy, w, h."""
    R, t, f, c = (np.asarray(a, dtype=np.float64) for a in (R, t, f, c))
    V = R.shape[0]
    wh = np.tile(np.asarray(image_size, dtype=np.float64), (V, 1))
    return np.concatenate([R.reshape(V, 9), t.reshape(V, 3), f.reshape(V, 2), c.reshape(V, 2), wh], axis=1)


def build_inputs(pix: torch.Tensor, calib) -> tuple:
    """pix [B, V, J, 3] fp32 (u, v, conf) pixels on the device; calib [V, 18] -> poses, rays [B,V,J,3], centers [B,V,1,3]."""
    B, V, J, _ = pix.shape
    device = pix.device
    pix = pix.to(torch.float32).contiguous()
    calib = torch.as_tensor(np.asarray(calib, dtype=np.float64)).to(device).contiguous()
    poses = torch.empty((B, V, J, 3), dtype=torch.float32, device=device)
    rays = torch.empty_like(poses)
    centers = torch.empty((B, V, 1, 3), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.lib().mpl_build_inputs(pix.data_ptr(), calib.data_ptr(), B, V, J, poses.data_ptr(), rays.data_ptr(),
                                               centers.data_ptr(), stream))
    return poses, rays, centers
