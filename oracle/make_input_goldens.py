"""ORACLE / test infrastructure — mint golden vectors for the input builder (N2) and the synthetic projector (N3) from the
UNMODIFIED reference code, in the build container (needs /root/reference; nothing here runs on the GPU box).

    python -m oracle.make_input_goldens            # writes tests/golden/inputs_{h36m,cmu}.npz

What is pinned, and by which reference code:

* N2 `mpl_build_inputs` / `oracle.build_inputs`: the reference's own `JointsDataset_MPL.__getitem__`
  (`MPL/lib/dataset/joints_dataset_mpl.py:443-811`) run on a synthetic db record per (pose, view) with the dataset flags of
  the shipped YAMLs (INPUTS_NORMALIZED, NORMALIZE_CAMERAS, USE_T, NO_AUGMENTATION, CLIP_JOINTS, OUTPUT_IN_METER,
  DOWNSAMPLE 1, USE_GRID false).  The instance is made with `object.__new__` (its `__init__` wants dataset files); every
  attribute `__getitem__` reads is set explicitly below.  Stored: the raw detections + calibrations going in, and the
  `joints` / `meta['rays']` / `meta['cam_center']` / `joints_3d` tensors coming out.
* N3 `mpl_synth_project`: `world_to_cam` + `cam_to_image` of `MPL/lib/utils/calib.py:42-77` applied to the 3D poses of the
  counter-based generator (`openmpl_b200/synth.py`), giving the raw pixels the device projector must reproduce.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

from . import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
LIB = os.path.join(ref_loader.REF_ROOT, "MPL/lib")


def load_dataset_module():
    """The reference `dataset.joints_dataset_mpl` module, loaded from its file unmodified (cv2 is installed here)."""
    if LIB not in sys.path:
        sys.path.insert(0, LIB)
    spec = importlib.util.spec_from_file_location("ref_joints_dataset_mpl", os.path.join(LIB, "dataset/joints_dataset_mpl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_calib_module():
    spec = importlib.util.spec_from_file_location("ref_utils_calib", os.path.join(LIB, "utils/calib.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_dataset(image_size, db):
    """A `JointsDataset_MPL` whose `__getitem__` runs on `db` with the dataset flags of the shipped YAMLs
    (configs/h36m/mpl_amass/hm_0_*.yaml:41-60, configs/cmu_panoptic/mpl_amass/cmu_0_*.yaml:8-27,61-64)."""
    mod = load_dataset_module()
    ds = object.__new__(mod.JointsDataset_MPL)
    ds.db = db
    ds.root, ds.data_format = "", "jpg"
    ds.is_train = False
    ds.num_joints = 17
    ds.image_size = np.array(image_size)
    ds.downsample = 1
    ds.use_grid = False
    ds.bug_test = False
    ds.use_t = True
    ds.output_in_meter = True
    ds.no_augmentation = True
    ds.clip_joints = True
    ds.inputs_normalized = True
    ds.normalize_cameras = True
    ds.flip_lower_body_kp = False
    ds.place_person_in_center = False
    ds.use_h36m_cameras_on_cmu = ds.use_cmu_cameras_on_h36m = ds.use_cmu_cameras_on_cmu = False
    ds.APPLY_NOISE = ds.APPLY_NOISE_MISSING = ds.APPLY_SMART_PSEUDO_TRAINING = False
    return ds


def mint(kind: str, views: int, poses: int, seed: int):
    """Run the reference on `poses` synthetic poses seen by `views` ring cameras of rig `kind`."""
    from openmpl_b200 import synth
    calib = load_calib_module()
    rig = synth.make_rig(views, kind)
    w, h = rig.image_size
    unit = 1000.0 if kind == "h36m" else 100.0          # the datasets store mm (H36M) / cm (CMU); OUTPUT_IN_METER divides
    source = "h36m" if kind == "h36m" else "cmu_panoptic"
    target = synth.make_batch(poses, rig, seed=seed)["target"].astype(np.float64)          # [B,17,3] metres, world
    rng = np.random.default_rng(seed + 1000)
    # spread the subjects so that a good part of the joints leave the image: exercises the clip / confidence rule
    target = target + rng.normal(0.0, 1.2, size=(poses, 1, 3)) * np.array([1.0, 1.0, 0.3])
    conf = rng.uniform(0.05, 1.0, size=(poses, views, 17))
    pix = np.zeros((poses, views, 17, 3))
    out_pose = np.zeros((poses, views, 17, 3), np.float32)
    out_rays = np.zeros((poses, views, 17, 3), np.float32)
    out_cent = np.zeros((poses, views, 1, 3), np.float32)
    out_3d = np.zeros((poses, views, 17, 3), np.float32)
    db = []
    for b in range(poses):
        for v in range(views):
            R = rig.R[v]
            t_ext = -R @ rig.t[v]                       # x_cam = R X + t_ext  (calib.py:42-59)
            K = np.array([[rig.f[v, 0], 0, rig.c[v, 0]], [0, rig.f[v, 1], rig.c[v, 1]], [0, 0, 1.0]])
            x_cam = calib.world_to_cam(target[b][None] * unit, R, (t_ext * unit).reshape(3, 1))[0]          # dataset units
            uv = calib.cam_to_image(x_cam[None], K)[0]                                                       # [17,2] pixels
            # the detections reach mpl_build_inputs as fp32: give the reference the same (rounded) numbers
            uv = uv.astype(np.float32).astype(np.float64)
            conf[b, v] = conf[b, v].astype(np.float32).astype(np.float64)
            pix[b, v, :, :2] = uv
            pix[b, v, :, 2] = conf[b, v]
            # With USE_T the dataset hands camera['t'] to the model as the "camera centre" and adds it to the rays
            # (joints_dataset_mpl.py:638-648,897-898).  The synthetic rigs put the camera POSITION there.
            cam = {"R": R.copy(), "T": (rig.t[v] * unit).reshape(3, 1).copy(), "t": (rig.t[v] * unit).reshape(3, 1).copy(),
                   "fx": float(rig.f[v, 0]), "fy": float(rig.f[v, 1]), "cx": float(rig.c[v, 0]), "cy": float(rig.c[v, 1]), "K": K}
            rec = {"source": source, "image": "none.jpg", "joints_2d": uv.copy(),
                   "joints_2d_conf": np.repeat(conf[b, v][:, None], 3, axis=1), "center": (w / 2.0, h / 2.0),
                   "scale": (w / 200.0, h / 200.0), "camera": cam, "subject": 0, "camera_id": v}
            if source == "h36m":
                # H36M records carry camera-frame joints; the dataset maps them back with cam_to_world(R, camera['t'])
                # (:528-529), i.e. it needs x_cam = R X + camera['t'] for THIS camera dict
                rec["joints_3d_camera"] = calib.world_to_cam(target[b][None] * unit, R, cam["t"])[0]
            else:
                rec["joints_3d"] = target[b] * unit
                rec["joints_3d_conf"] = np.ones(17)
            db.append(rec)
    ds = reference_dataset((w, h), db)
    for b in range(poses):
        for v in range(views):
            joints, joints_3d, _, meta = ds[b * views + v]
            out_pose[b, v] = joints.numpy()
            out_rays[b, v] = meta["rays"].numpy()
            out_cent[b, v] = meta["cam_center"].numpy()
            out_3d[b, v] = joints_3d.numpy()
    calib18 = np.concatenate([rig.R.reshape(views, 9), rig.t, rig.f, rig.c, np.tile([float(w), float(h)], (views, 1))], axis=1)
    return {"pix": pix.astype(np.float32), "calib": calib18, "target": target.astype(np.float32),
            "poses": out_pose, "rays": out_rays, "centers": out_cent, "joints_3d": out_3d,
            "kind": np.array(kind), "seed": np.array(seed)}


def mint_projection(kind: str, views: int, poses: int, seed: int, start: int):
    """calib.world_to_cam / cam_to_image (calib.py:42-77) on the generator's own 3D poses: what mpl_synth_project must give."""
    from openmpl_b200 import synth
    calib = load_calib_module()
    rig = synth.make_rig(views, kind)
    target = synth.make_batch(poses, rig, seed=seed, start=start)["target"].astype(np.float64)
    uv = np.zeros((poses, views, 17, 2))
    for v in range(views):
        K = np.array([[rig.f[v, 0], 0, rig.c[v, 0]], [0, rig.f[v, 1], rig.c[v, 1]], [0, 0, 1.0]])
        x_cam = calib.world_to_cam(target, rig.R[v], (-rig.R[v] @ rig.t[v]).reshape(3, 1))
        uv[:, v] = calib.cam_to_image(x_cam, K)
    return {"proj_uv": uv, "proj_target": target.astype(np.float32), "proj_seed": np.array(seed), "proj_start": np.array(start),
            "proj_views": np.array(views)}


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for kind, views in (("h36m", 4), ("cmu", 5)):
        g = mint(kind, views, poses=24, seed=5)
        g.update(mint_projection(kind, views, poses=64, seed=3, start=1000))
        path = os.path.join(GOLDEN_DIR, f"inputs_{kind}.npz")
        np.savez_compressed(path, **g)
        clipped = float((g["poses"][..., 2] == 0).mean())
        print(f"{path}: {os.path.getsize(path)} bytes, {clipped:.0%} of the joints clipped / zero confidence")


if __name__ == "__main__":
    main()
