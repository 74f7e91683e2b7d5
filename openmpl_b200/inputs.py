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
    if isinstance(calib, torch.Tensor):
        calib = calib.to(device=device, dtype=torch.float64).contiguous()
    else:
        calib = torch.as_tensor(np.asarray(calib, dtype=np.float64)).to(device).contiguous()
    poses = torch.empty((B, V, J, 3), dtype=torch.float32, device=device)
    rays = torch.empty_like(poses)
    centers = torch.empty((B, V, 1, 3), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.lib().mpl_build_inputs(pix.data_ptr(), calib.data_ptr(), B, V, J, poses.data_ptr(), rays.data_ptr(),
                                               centers.data_ptr(), stream))
    return poses, rays, centers


def synth_project(batch: int, rig, seed: int = 0, start: int = 0, conf_mode: str = "uniform", device=None) -> tuple:
    """N3: MHP-style synthetic poses generated and projected ON THE DEVICE (`mpl_synth_project`).

    Pose i depends only on (seed, start + i) — the same Philox stream as `synth.make_batch` — so ranks / micro-batches
    can shard the global index range freely.  Returns (pix [B,V,17,3] raw pixels (u, v, conf), target [B,17,3] metres,
    calib [V,18] fp64 on the device); `build_inputs(pix, calib)` turns the pixels into the model's inputs.
    """
    from .synth import NUM_JOINTS
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    V, J = rig.num_views, NUM_JOINTS
    calib = torch.as_tensor(pack_calibration(rig.R, rig.t, rig.f, rig.c, rig.image_size)).to(device)
    room = torch.as_tensor(np.asarray(rig.room, dtype=np.float64)).to(device)
    pix = torch.empty((batch, V, J, 3), dtype=torch.float32, device=device)
    target = torch.empty((batch, J, 3), dtype=torch.float32, device=device)
    if conf_mode not in ("uniform", "ones"):
        raise ValueError(conf_mode)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.lib().mpl_synth_project(int(seed), int(start), int(batch), V, J, calib.data_ptr(), room.data_ptr(),
                                                int(conf_mode == "ones"), pix.data_ptr(), target.data_ptr(), stream))
    return pix, target, calib
