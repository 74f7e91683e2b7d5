"""K6 (MPJPE accumulators) and N2 (input builder) kernels against the oracle restatements."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

from oracle import mpl_oracle
from openmpl_b200 import inputs, metric, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,J,masked", [(1, 17, False), (50, 17, False), (50, 17, True), (4097, 17, True), (33, 40, True), (0, 17, False)])
def test_mpjpe_accumulator_matches_reference_metric(B, J, masked):
    rng = np.random.default_rng(B + J)
    pred = rng.normal(size=(B, J, 3)).astype(np.float32)
    gt = rng.normal(size=(B, J, 3)).astype(np.float32)
    conf = None
    if masked:
        conf = (rng.random((B, J, 1)) > 0.2).astype(np.float32)
        if B > 3:
            conf[3, 0] = 0            # a masked root joint: blanks the pose in the root-relative metric
    acc = metric.MpjpeAccumulator(J, output_in_meter=True)
    half = B // 2                      # two updates must equal one: the sums are running
    for sl in (slice(0, half), slice(half, B)):
        acc.update(torch.from_numpy(pred[sl]).cuda(), torch.from_numpy(gt[sl]).cuda(),
                   None if conf is None else torch.from_numpy(conf[sl]).cuda())
    got = acc.acc.cpu().numpy()
    if B == 0:
        assert np.all(got == 0)
        return
    want = mpl_oracle.metric_sums(pred, gt, conf)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
    res = acc.result()
    a = mpl_oracle.evaluate(pred, gt, True, None if conf is None else np.broadcast_to(conf, (B, J, 3)), relative=False)
    r = mpl_oracle.evaluate(pred, gt, True, None if conf is None else np.broadcast_to(conf, (B, J, 3)), relative=True)
    np.testing.assert_allclose(res["pjpe_abs"], a["pjpe"], rtol=1e-9)
    np.testing.assert_allclose(res["mpjpe_rel"], r["mpjpe"], rtol=1e-9)
    np.testing.assert_allclose(np.nan_to_num(res["dist_abs"]), np.nan_to_num(a["dist_per_dim_per_kp"]), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.nan_to_num(res["dist_rel"]), np.nan_to_num(r["dist_per_dim_per_kp"]), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("equal", [True, False])
def test_mpjpe_accumulator_room_unscaling_and_3d_confidence(equal):
    """The tail of validate() for room-normalised datasets (function_mpl.py:476-490): predictions and targets are un-scaled
    (v * s + centre, or x / y scaled separately) before the metric, joints with joints_3d_conf <= 0 are masked."""
    rng = np.random.default_rng(11)
    B, J = 300, 17
    pred = rng.normal(size=(B, J, 3)).astype(np.float32)
    gt = (pred + rng.normal(0, 0.05, size=(B, J, 3))).astype(np.float32)
    conf = (rng.random((B, J)) > 0.1).astype(np.float32)
    room = {"room_x_scale": 2.5, "room_center": [0.25, -0.5, 0.9]} if equal else {"room_x_scale": 2.5, "room_y_scale": 1.75}
    acc = metric.MpjpeAccumulator(J, output_in_meter=True)
    acc.update(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(conf).cuda(), room=room)
    p2, g2 = mpl_oracle.room_unscale(pred, gt, room)
    want = mpl_oracle.metric_sums(p2, g2, conf[:, :, None])
    np.testing.assert_allclose(acc.acc.cpu().numpy(), want, rtol=1e-9, atol=1e-9)
    plain = metric.MpjpeAccumulator(J, output_in_meter=True)
    plain.update(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(conf).cuda())
    assert abs(plain.result()["mpjpe_abs"] - acc.result()["mpjpe_abs"]) > 1e-3       # the un-scaling changes the metric


PROCRUSTES_MODES = [(True, "best"), (False, "best"), (True, False), (True, True)]


@pytest.mark.parametrize("scaling,reflection", PROCRUSTES_MODES)
def test_pmpjpe_accumulator_matches_reference_procrustes_golden(scaling, reflection):
    """Device P-MPJPE sums against PoseUtils.procrustes outputs of the unmodified reference (tests/golden/procrustes.npz)."""
    g = np.load(os.path.join(GOLDEN_DIR, "procrustes.npz"))
    A, B = g["A"], g["B"]                                   # gt, prediction
    tag = f"s{int(scaling)}_r{reflection}"
    J = A.shape[1]
    acc = metric.PmpjpeAccumulator(J, output_in_meter=False, scaling=scaling, reflection=reflection)
    acc.update(torch.from_numpy(B).cuda(), torch.from_numpy(A).cuda())
    got = acc.acc.cpu().numpy()
    want = np.concatenate([np.sqrt(((g["Z_" + tag] - A.astype(np.float64)) ** 2).sum(-1)).sum(0),
                           [g["d_" + tag].sum(), g["scale_" + tag].sum(), len(A)]])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("B,J", [(1, 17), (63, 17), (64, 17), (65, 17), (4097, 17), (200, 13), (70, 40), (0, 17)])
def test_pmpjpe_accumulator_matches_oracle(B, J):
    rng = np.random.default_rng(B * 7 + J)
    gt = (rng.normal(0, 0.3, size=(B, J, 3)) + rng.uniform(-2, 2, size=(B, 1, 3))).astype(np.float32)
    pred = (gt + rng.normal(0, 0.05, size=(B, J, 3))).astype(np.float32)
    if B > 5:
        pred[5, :, 2] = 0.25                                # a planar prediction: one singular value vanishes
    acc = metric.PmpjpeAccumulator(J, output_in_meter=True)
    half = B // 2
    for sl in (slice(0, half), slice(half, B)):
        acc.update(torch.from_numpy(pred[sl]).cuda(), torch.from_numpy(gt[sl]).cuda())
    got = acc.acc.cpu().numpy()
    if B == 0:
        assert np.all(got == 0)
        return
    want = mpl_oracle.pmpjpe_sums(pred, gt, output_in_meter=True)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-8)
    res = acc.result()
    assert res["n"] == B and abs(res["p_mpjpe"] - want[:J].mean() / B) < 1e-9


def test_pmpjpe_properties_at_full_batch():
    """B = 65 536 (the bench batch): similarity-transformed ground truth scores ~0; the metric is invariant under a
    similarity transform of the predictions; rigid alignment (scaling off) of a rigidly moved pose scores ~0 too."""
    B, J = 65536, 17
    gen = torch.Generator(device="cuda").manual_seed(4)
    gt = torch.randn(B, J, 3, device="cuda", generator=gen) * 0.3 + torch.rand(B, 1, 3, device="cuda", generator=gen) * 4 - 2
    Q = torch.linalg.qr(torch.randn(3, 3, device="cuda", generator=gen, dtype=torch.float64))[0].float()
    moved = 1.3 * gt @ Q + torch.tensor([0.5, -1.0, 2.0], device="cuda")
    a = metric.PmpjpeAccumulator(J, output_in_meter=True)
    a.update(moved, gt)
    r = a.result()
    assert r["n"] == B and r["p_mpjpe"] < 2e-4 and abs(r["scale"] - 1 / 1.3) < 1e-5      # cm; fp32 input rounding only
    rigid = metric.PmpjpeAccumulator(J, output_in_meter=True, scaling=False)
    rigid.update(gt @ Q + 0.7, gt)
    assert rigid.result()["p_mpjpe"] < 2e-4
    pred = gt + 0.03 * torch.randn(B, J, 3, device="cuda", generator=gen)
    p0 = metric.PmpjpeAccumulator(J)
    p0.update(pred, gt)
    p1 = metric.PmpjpeAccumulator(J)
    p1.update(0.6 * pred @ Q - 1.5, gt)
    assert abs(p0.result()["p_mpjpe"] - p1.result()["p_mpjpe"]) < 2e-4
    plain = metric.MpjpeAccumulator(J)
    plain.update(pred, gt)
    assert p0.result()["p_mpjpe"] < plain.result()["mpjpe_abs"]


@pytest.mark.parametrize("kind,V,B", [("h36m", 4, 64), ("cmu", 5, 257), ("h36m", 8, 3), ("cmu", 2, 0)])
def test_input_builder_matches_dataset_math(kind, V, B):
    rig = synth.make_rig(V, kind)
    rng = np.random.default_rng(V * 100 + B)
    w, h = rig.image_size
    pix = np.concatenate([rng.uniform(-0.15 * w, 1.15 * w, (B, V, 17, 1)), rng.uniform(-0.15 * h, 1.15 * h, (B, V, 17, 1)),
                          rng.uniform(0.05, 1.0, (B, V, 17, 1))], axis=-1).astype(np.float32)
    if B:
        pix[0, 0, 0, :2] = (0.0, 5.0)            # exactly on the border: "0 < x" is false -> confidence zeroed
        pix[0, 0, 1, :2] = (w - 1.0, 5.0)
    want = mpl_oracle.build_inputs(pix, rig.R, rig.t, rig.f, rig.c, rig.image_size)
    calib = inputs.pack_calibration(rig.R, rig.t, rig.f, rig.c, rig.image_size)
    got = inputs.build_inputs(torch.from_numpy(pix).cuda(), calib)
    torch.cuda.synchronize()
    for g, wv in zip(got, want):
        np.testing.assert_allclose(g.cpu().numpy(), wv, rtol=0, atol=2e-7 * max(1.0, float(np.abs(wv).max()) if wv.size else 1.0))
    if B:
        assert float(got[0][0, 0, 0, 2]) == 0.0 and float(got[0][0, 0, 1, 2]) == 0.0
        assert (got[0][..., 2] == 0).any() and (got[0][..., 2] > 0).any()


def test_input_builder_reproduces_the_synthetic_generator():
    """synth.make_batch (3D -> pixels -> inputs) and the kernel (pixels -> inputs) agree when fed the same pixels."""
    rig = synth.make_rig(4, "h36m")
    batch = synth.make_batch(200, rig, seed=5)
    w, h = rig.image_size
    # invert the screen normalisation to recover the (already clipped) pixels
    px = (batch["poses"][..., :2].astype(np.float64) + np.array([1, h / w])) / 2 * w
    pix = np.concatenate([px, batch["poses"][..., 2:3]], axis=-1).astype(np.float32)
    got = inputs.build_inputs(torch.from_numpy(pix).cuda(), inputs.pack_calibration(rig.R, rig.t, rig.f, rig.c, rig.image_size))
    np.testing.assert_allclose(got[1].cpu().numpy(), batch["rays"], atol=2e-3)     # pixel round-trip through fp32
    np.testing.assert_allclose(got[2].cpu().numpy(), batch["centers"], atol=1e-6)


@pytest.mark.parametrize("kind,V", [("h36m", 4), ("cmu", 5), ("h36m", 8)])
def test_device_projector_reproduces_the_host_generator(kind, V):
    """N3: the on-device Philox generator + projector + input builder gives the tensors synth.make_batch builds on the
    host (same uniforms bit for bit; fp64 libm differences and the fp32 pixel round trip stay below 2e-6)."""
    from openmpl_b200 import inputs, synth
    rig = synth.make_rig(V, kind)
    B, start = 257, 1000
    ref = synth.make_batch(B, rig, seed=3, start=start)
    pix, target, calib = inputs.synth_project(B, rig, seed=3, start=start)
    poses, rays, centers = inputs.build_inputs(pix, calib)
    torch.cuda.synchronize()
    np.testing.assert_allclose(target.cpu().numpy(), ref["target"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(poses.cpu().numpy(), ref["poses"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(rays.cpu().numpy(), ref["rays"], rtol=0, atol=5e-6)
    np.testing.assert_array_equal(centers.cpu().numpy(), ref["centers"])
    # sharding invariance: any split of the global index range yields the same poses
    a, ta, _ = inputs.synth_project(100, rig, seed=3, start=start)
    b, tb, _ = inputs.synth_project(157, rig, seed=3, start=start + 100)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(torch.cat([a, b]).cpu().numpy(), pix.cpu().numpy())
    np.testing.assert_array_equal(torch.cat([ta, tb]).cpu().numpy(), target.cpu().numpy())


def test_device_evaluation_loop_matches_host_oracle_metric():
    """Config-5 style loop (device generator -> input builder -> forward -> fp64 MPJPE sums, micro-batched) against the
    oracle forward + the reference's evaluate() on the host generator's data."""
    from openmpl_b200 import evaluate, spec, synth
    from oracle import mpl_oracle
    n = 300
    line = evaluate.run(arch="cmu0", views=2, poses=n, micro_batch=128, precision="fp32", seed=4)
    kw = dict(num_joints=17, embed_dim_ratio=32, num_heads=8, depth=2, num_views=2, drop_path_rate=0.1, **spec.HM0_FLAGS)
    cfg = spec.make_config(**kw)
    weights = synth.named_weights(spec.param_spec(cfg), seed=0)
    batch = synth.make_batch(n, synth.make_rig(2, "cmu"), seed=4)
    ref = mpl_oracle.forward(weights, cfg, batch["poses"], batch["rays"], batch["centers"])
    ev_a = mpl_oracle.evaluate(ref, batch["target"], output_in_meter=True, relative=False)
    ev_r = mpl_oracle.evaluate(ref, batch["target"], output_in_meter=True, relative=True)
    assert line["poses"] == n
    assert abs(line["mpjpe_cm"]["absolute"] - ev_a["mpjpe"]) < 2e-3       # cm; inputs differ by <= 2e-6 from the host generator
    assert abs(line["mpjpe_cm"]["root_relative"] - ev_r["mpjpe"]) < 2e-3
    pm = mpl_oracle.pmpjpe_sums(ref, batch["target"], output_in_meter=True)
    assert abs(line["mpjpe_cm"]["procrustes_aligned"] - pm[:17].mean() / n) < 2e-3
