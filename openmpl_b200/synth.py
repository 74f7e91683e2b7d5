"""Synthetic multi-view inputs for the MPL lifter forward.

Produces exactly the tensors the reference dataset hands to the model
(`MPL/lib/dataset/joints_dataset_mpl.py:443-811`), from synthetic 3D poses and
synthetic ring-camera calibrations:

* 3D pose: root uniform in the YAML room, other joints root + N(0, 0.25 m)
* projection `x_cam = R X + T_ext`, `u = K x_cam / z`   (`MPL/lib/utils/calib.py:42-77`)
* clip to the image and zero the confidence of clipped joints (`joints_dataset_mpl.py:701-715`)
* screen normalisation `(u / w) * 2 - [1, h / w]`          (`joints_dataset_mpl.py:817-820`)
* intrinsics normalised the same way                      (`joints_dataset_mpl.py:615-623`)
* rays `R^T [(x-cx)/fx, (y-cy)/fy, 1] + t`, centers `t^T` (`joints_dataset_mpl.py:638-648,872-898`)

The generator is counter based: pose `i` of seed `s` is the same whatever the
batch size or the rank that asks for it (Philox counter = i * blocks-per-pose),
so sharding over ranks yields identical data to a single-GPU run.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

NUM_JOINTS = 17


@dataclass(frozen=True)
class Rig:
    """V ring cameras looking at the room centre (all angles in radians, lengths in metres)."""
    R: np.ndarray          # [V,3,3] world -> camera rotation
    t: np.ndarray          # [V,3]   camera position in the world ("USE_T" convention: the model's `centers`)
    f: np.ndarray          # [V,2]   focal length, pixels
    c: np.ndarray          # [V,2]   principal point, pixels
    image_size: tuple      # (w, h) pixels
    room: tuple            # (min_x, max_x, min_y, max_y) metres

    @property
    def num_views(self) -> int:
        return self.R.shape[0]


def make_rig(num_views: int, kind: str = "h36m") -> Rig:
    """Ring of cameras, radius 4.5 m, height 1.5 m, looking at the room centre.

    kind "h36m": f=1145 px, 1000x1000 image, room -1..1 x -1.5..2  (hm_0 yaml:47-50,89-91)
    kind "cmu" : f=1400 px, 1920x1080 image, room -2..0.3 x -0.8..0.8 (cmu_0 yaml:85-87)
    """
    if kind == "h36m":
        f0, (w, h), room = 1145.0, (1000, 1000), (-1.0, 1.0, -1.5, 2.0)
    elif kind == "cmu":
        f0, (w, h), room = 1400.0, (1920, 1080), (-2.0, 0.3, -0.8, 0.8)
    else:
        raise ValueError(f"unknown rig kind {kind!r}")
    centre = np.array([(room[0] + room[1]) / 2, (room[2] + room[3]) / 2, 0.9])
    Rs, ts = [], []
    for v in range(num_views):
        ang = 2 * math.pi * v / num_views + 0.3
        pos = np.array([centre[0] + 4.5 * math.cos(ang), centre[1] + 4.5 * math.sin(ang), 1.5])
        z = centre - pos
        z /= np.linalg.norm(z)                       # optical axis
        x = np.cross(z, np.array([0.0, 0.0, 1.0]))
        x /= np.linalg.norm(x)
        y = np.cross(z, x)                           # image y points down
        Rs.append(np.stack([x, y, z]))               # rows = camera axes in world coords
        ts.append(pos)
    V = num_views
    return Rig(R=np.stack(Rs), t=np.stack(ts), f=np.full((V, 2), f0), c=np.tile([w / 2.0, h / 2.0], (V, 1)),
               image_size=(w, h), room=room)


def _uniforms(seed: int, start: int, count: int, per_pose: int) -> np.ndarray:
    """[count, per_pose] uniforms in [0,1); row i depends only on (seed, start+i)."""
    blocks = (per_pose + 3) // 4                      # Philox4x64: 4 outputs per counter step
    bg = np.random.Philox(key=seed)
    bg.advance(start * blocks)
    raw = bg.random_raw(count * blocks * 4).reshape(count, blocks * 4)[:, :per_pose]
    return (raw >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def make_batch(batch: int, rig: Rig, seed: int = 0, start: int = 0, conf_mode: str = "uniform",
               dtype=np.float32) -> dict:
    """Model inputs for poses [start, start+batch).

    Returns dict with `poses [B,V,17,3]` (x̂, ŷ, conf), `rays [B,V,17,3]`, `centers [B,V,1,3]`,
    `target [B,17,3]` (metres, OUTPUT_IN_METER convention) — all `dtype`.
    """
    V, J = rig.num_views, NUM_JOINTS
    per_pose = 3 + 2 * 3 * (J - 1) + V * J
    u = _uniforms(seed, start, batch, per_pose)
    room = rig.room
    root = np.stack([room[0] + u[:, 0] * (room[1] - room[0]),
                     room[2] + u[:, 1] * (room[3] - room[2]),
                     0.8 + 0.2 * u[:, 2]], axis=1)                         # [B,3]
    n = 3 * (J - 1)
    u1 = np.maximum(u[:, 3:3 + n], 1e-12)
    u2 = u[:, 3 + n:3 + 2 * n]
    gauss = np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * math.pi * u2)          # Box-Muller, fixed draw count
    target = np.concatenate([root[:, None, :], root[:, None, :] + 0.25 * gauss.reshape(batch, J - 1, 3)], axis=1)
    conf_u = u[:, 3 + 2 * n:].reshape(batch, V, J)

    w, h = rig.image_size
    # world -> camera -> pixels (calib.py:42-77), per view
    x_cam = np.einsum("vij,bkj->bvki", rig.R, target) - np.einsum("vij,vj->vi", rig.R, rig.t)[None, :, None, :]
    z = x_cam[..., 2:3]
    px = x_cam[..., :2] / z * rig.f[None, :, None, :] + rig.c[None, :, None, :]      # [B,V,J,2]
    # clip + confidence zeroing (joints_dataset_mpl.py:701-715)
    inside = (px[..., 0] > 0) & (px[..., 0] < w - 1) & (px[..., 1] > 0) & (px[..., 1] < h - 1) & (z[..., 0] > 0)
    if conf_mode == "uniform":
        conf = 0.3 + 0.7 * conf_u
    elif conf_mode == "ones":
        conf = np.ones_like(conf_u)
    else:
        raise ValueError(conf_mode)
    conf = np.where(inside, conf, 0.0)
    px = np.stack([np.clip(px[..., 0], 0, w - 1), np.clip(px[..., 1], 0, h - 1)], axis=-1)
    # screen normalisation (joints_dataset_mpl.py:817-820) of joints and intrinsics (:615-623)
    norm = np.array([1.0, h / w])
    xy = px / w * 2 - norm
    c_hat = rig.c / w * 2 - norm                                            # [V,2]
    f_hat = rig.f / w * 2
    # rays (joints_dataset_mpl.py:872-898) with USE_T: R^T [x, y, 1] + t ; centers = t^T (:645-646)
    cam_dir = np.concatenate([(xy - c_hat[None, :, None, :]) / f_hat[None, :, None, :],
                              np.ones((batch, V, J, 1))], axis=-1)
    rays = np.einsum("vji,bvkj->bvki", rig.R, cam_dir) + rig.t[None, :, None, :]
    centers = np.broadcast_to(rig.t[None, :, None, :], (batch, V, 1, 3))
    poses = np.concatenate([xy, conf[..., None]], axis=-1)
    return {"poses": np.ascontiguousarray(poses, dtype=dtype), "rays": np.ascontiguousarray(rays, dtype=dtype),
            "centers": np.ascontiguousarray(centers, dtype=dtype), "target": np.ascontiguousarray(target, dtype=dtype)}


def named_weights(shapes: dict, seed: int = 0, pos_std: float = 0.02) -> dict:
    """Deterministic weights keyed by parameter NAME (order independent, numpy only).

    Distributions follow the PyTorch defaults the reference relies on:
    Linear/Conv1d weight and bias ~ U(±1/sqrt(fan_in)); LayerNorm/BatchNorm weight 1 + small noise,
    bias small noise (so the affine paths are exercised); learned position embeddings ~ N(0, pos_std²)
    (zeros at reference init — randomised so the add paths are tested); BN running stats non-trivial.
    `shapes`: name -> (shape tuple, kind) with kind in {linear_w, linear_b, norm_w, norm_b, pos, bn_mean, bn_var, count}.
    """
    import zlib
    out = {}
    for name, (shape, kind, fan_in) in shapes.items():
        key = (zlib.crc32(name.encode()) << 20) ^ (seed & 0xFFFFF)
        rng = np.random.Generator(np.random.Philox(key=key))
        size = int(np.prod(shape)) if len(shape) else 1
        if kind in ("linear_w", "linear_b"):
            bound = 1.0 / math.sqrt(fan_in)
            a = (rng.random(size) * 2 - 1) * bound
        elif kind == "norm_w":
            a = 1.0 + 0.1 * (rng.random(size) * 2 - 1)
        elif kind == "norm_b":
            a = 0.05 * (rng.random(size) * 2 - 1)
        elif kind == "pos":
            a = pos_std * rng.standard_normal(size)
        elif kind == "bn_mean":
            a = 0.1 * (rng.random(size) * 2 - 1)
        elif kind == "bn_var":
            a = 0.5 + rng.random(size)
        elif kind == "count":
            out[name] = np.zeros(shape, dtype=np.int64)
            continue
        else:
            raise ValueError(kind)
        out[name] = a.astype(np.float32).reshape(shape)
    return out
