"""N2 / N3 parity pins.  tests/golden/inputs_{h36m,cmu}.npz were minted by oracle/make_input_goldens.py from the UNMODIFIED
reference: `JointsDataset_MPL.__getitem__` (MPL/lib/dataset/joints_dataset_mpl.py:443-811) run on synthetic db records with
the dataset flags of the shipped YAMLs, and `world_to_cam` / `cam_to_image` (MPL/lib/utils/calib.py:42-77).

CPU part: the oracle restatement against those goldens (bit-exact) and, when /root/reference is mounted, a live re-mint
against the committed files.  GPU part: `mpl_build_inputs` and `mpl_synth_project` against the same goldens."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import mpl_oracle, ref_loader

KINDS = ["h36m", "cmu"]


def _golden(kind):
    return np.load(os.path.join(GOLDEN_DIR, f"inputs_{kind}.npz"))


def _calib(g):
    c = g["calib"]
    V = c.shape[0]
    return c[:, :9].reshape(V, 3, 3), c[:, 9:12], c[:, 12:14], c[:, 14:16], (c[0, 16], c[0, 17])


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_input_builder_is_the_reference_getitem(kind):
    g = _golden(kind)
    R, t, f, c, size = _calib(g)
    poses, rays, centers = mpl_oracle.build_inputs(g["pix"], R, t, f, c, size)
    np.testing.assert_array_equal(poses, g["poses"])          # x^, y^, confidence (clip + zeroing rule :709-715)
    np.testing.assert_array_equal(rays, g["rays"])            # R^T [(x-cx)/fx, (y-cy)/fy, 1] + t   (:872-898)
    np.testing.assert_array_equal(centers, g["centers"])      # camera['t']^T with USE_T            (:645-646)
    # the goldens exercise both branches of the clip rule
    conf = g["poses"][..., 2]
    assert (conf == 0).any() and (conf > 0).any()
    # OUTPUT_IN_METER: the dataset's 3D target is the world pose in metres for every view (:513-576)
    np.testing.assert_allclose(g["joints_3d"], np.broadcast_to(g["target"][:, None], g["joints_3d"].shape), atol=1e-6)


@pytest.mark.parametrize("kind", KINDS)
def test_host_generator_projection_is_the_reference_calib(kind):
    """synth.make_batch projects with x_cam = R (X - pos), u = f x / z + c; the reference's world_to_cam / cam_to_image
    on the same 3D poses must give the same pixels (up to fp64 round-off of the different operation order)."""
    from openmpl_b200 import synth
    g = _golden(kind)
    V = int(g["proj_views"])
    rig = synth.make_rig(V, kind)
    B = g["proj_uv"].shape[0]
    target = synth.make_batch(B, rig, seed=int(g["proj_seed"]), start=int(g["proj_start"]))["target"]
    np.testing.assert_array_equal(target, g["proj_target"])
    t64 = target.astype(np.float64)
    x_cam = np.einsum("vij,bkj->bvki", rig.R, t64) - np.einsum("vij,vj->vi", rig.R, rig.t)[None, :, None, :]
    uv = x_cam[..., :2] / x_cam[..., 2:3] * rig.f[None, :, None, :] + rig.c[None, :, None, :]
    np.testing.assert_allclose(uv, g["proj_uv"], rtol=0, atol=1e-9)


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted (GPU box)")
@pytest.mark.parametrize("kind,views", [("h36m", 4), ("cmu", 5)])
def test_committed_input_goldens_are_what_the_reference_produces(kind, views):
    from oracle import make_input_goldens as mk
    live = mk.mint(kind, views, poses=24, seed=5)
    live.update(mk.mint_projection(kind, views, poses=64, seed=3, start=1000))
    g = _golden(kind)
    for k in ("pix", "poses", "rays", "centers", "joints_3d", "proj_uv"):
        np.testing.assert_array_equal(live[k], g[k], err_msg=k)


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted (GPU box)")
def test_reference_helpers_called_unbound():
    """`normalize_screen_coordinates` (:817-820) and `create_3d_ray_coords` (:872-904) of the unmodified class, called on a
    bare instance, against the oracle."""
    from types import SimpleNamespace
    from oracle import make_input_goldens as mk
    cls = mk.load_dataset_module().JointsDataset_MPL
    X = np.array([[0.0, 0.0], [999.0, 999.0], [500.0, 250.0]])
    np.testing.assert_array_equal(cls.normalize_screen_coordinates(None, X.copy(), 1000, 1000),
                                  mpl_oracle.normalize_screen_coordinates(X.copy(), 1000, 1000))
    np.testing.assert_array_equal(cls.normalize_screen_coordinates(None, X.copy(), 1920, 1080),
                                  mpl_oracle.normalize_screen_coordinates(X.copy(), 1920, 1080))
    rng = np.random.default_rng(0)
    Rm = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    cam = {"R": Rm, "t": rng.normal(size=(3, 1)), "T": rng.normal(size=(3, 1)), "fx": 2.29, "fy": 2.29, "cx": 0.01, "cy": -0.02}
    me = SimpleNamespace(downsample=1, use_grid=False, use_t=True, bug_test=False)
    joints = rng.uniform(-1, 1, size=(17, 2))
    rays = cls.create_3d_ray_coords(me, cam, None, joints.copy()).numpy()
    d = np.concatenate([(joints - [cam["cx"], cam["cy"]]) / [cam["fx"], cam["fy"]], np.ones((17, 1))], axis=1)
    want = (d @ Rm + cam["t"].T).astype(np.float32)
    np.testing.assert_allclose(rays, want, rtol=0, atol=1e-6)


# ---- GPU: the kernels against the same goldens ------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_device_input_builder_matches_the_reference_getitem(kind):
    import torch
    from openmpl_b200 import inputs
    g = _golden(kind)
    got = inputs.build_inputs(torch.from_numpy(g["pix"]).cuda(), g["calib"])
    torch.cuda.synchronize()
    poses, rays, centers = (x.cpu().numpy() for x in got)
    # the kernel evaluates the same fp64 expressions; only the fused-multiply-add contraction of the 3x3 product may differ
    np.testing.assert_array_equal(poses, g["poses"])
    np.testing.assert_array_equal(centers, g["centers"])
    np.testing.assert_allclose(rays, g["rays"], rtol=0, atol=1e-6)
    assert (rays != g["rays"]).mean() < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_device_projector_matches_the_reference_calib(kind):
    """`mpl_synth_project` pixels against calib.world_to_cam + cam_to_image of the reference on the same 3D poses."""
    import torch
    from openmpl_b200 import inputs, synth
    g = _golden(kind)
    V = int(g["proj_views"])
    rig = synth.make_rig(V, kind)
    B = g["proj_uv"].shape[0]
    pix, target, _ = inputs.synth_project(B, rig, seed=int(g["proj_seed"]), start=int(g["proj_start"]))
    torch.cuda.synchronize()
    np.testing.assert_allclose(target.cpu().numpy(), g["proj_target"], rtol=0, atol=2e-6)   # libm log / cos differ in the last bit
    uv = pix.cpu().numpy()[..., :2].astype(np.float64)
    # pixels are stored as fp32 (half an ulp at 2000 px = 6e-5) on top of the 2e-6 m difference of the 3D points
    np.testing.assert_allclose(uv, g["proj_uv"], rtol=0, atol=2e-3)
    conf = pix.cpu().numpy()[..., 2]
    assert ((conf >= 0.3) & (conf <= 1.0)).all()
