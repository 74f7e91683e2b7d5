"""The oracle (numpy restatement) against the goldens minted from the unmodified reference, and — when
/root/reference is mounted — against the live reference module. CPU only."""
import itertools
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR as GOLDEN, load_golden
from oracle import mpl_oracle, ref_loader
from oracle.cases import CASES, GRID_FLAGS, GRID_BASE, make_inputs
from openmpl_b200 import spec

SMALL = [n for n in CASES if not n.endswith("_d12")]
FULL = [n for n in CASES if n.endswith("_d12")]


def _check(name):
    case = CASES[name]
    g = load_golden(name)
    cfg, weights, batch = make_inputs(case)
    # the regenerated inputs are the stored ones (generator determinism across machines)
    for k in ("poses", "rays", "centers"):
        np.testing.assert_array_equal(batch[k], g[k])
    out = mpl_oracle.forward(weights, cfg, batch["poses"], batch["rays"], batch["centers"])
    outs = [out[0]] + list(out[1]) if isinstance(out, tuple) else [out]
    for i, o in enumerate(outs):
        scale = np.abs(g[f"out64_{i}"]).max()
        assert np.abs(o - g[f"out64_{i}"]).max() <= 1e-12 * max(scale, 1.0), name       # fp64 vs fp64 reference
        assert np.abs(o - g[f"out32_{i}"]).max() <= 2e-5 * max(scale, 1.0), name        # fp32 reference round-off


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_golden_small(name):
    _check(name)


@pytest.mark.parametrize("name", FULL)
def test_oracle_matches_golden_full_arch(name):
    _check(name)


def test_oracle_fp32_mode_close():
    cfg, weights, batch = make_inputs(CASES["flag_hm0flags_small"])
    a = mpl_oracle.forward(weights, cfg, batch["poses"], batch["rays"], batch["centers"], dtype=np.float32)
    b = mpl_oracle.forward(weights, cfg, batch["poses"], batch["rays"], batch["centers"])
    assert a.dtype == np.float32 and np.abs(a - b).max() < 1e-4


def test_validity_matches_reference_grid():
    """spec.make_config().error reproduces exactly which of the 4096 flag combinations the reference runs."""
    g = load_golden("validity_grid")
    assert g["meta"]["flags"] == GRID_FLAGS and g["meta"]["base"] == GRID_BASE
    ok = g["ok"]
    assert int(ok.sum()) == 1776
    for idx, bits in enumerate(itertools.product((False, True), repeat=len(GRID_FLAGS))):
        cfg = spec.make_config(**dict(GRID_BASE, **dict(zip(GRID_FLAGS, bits))))
        assert (cfg.error is None) == bool(ok[idx]), (bits, cfg.error)


def test_oracle_runs_every_valid_grid_combo_sampled():
    """The oracle executes (shape-wise) on a sample of valid combinations and raises on invalid ones."""
    g = load_golden("validity_grid")
    rng = np.random.default_rng(0)
    from openmpl_b200 import synth
    rig = synth.make_rig(GRID_BASE["num_views"])
    batch = synth.make_batch(2, rig, seed=5)
    combos = list(itertools.product((False, True), repeat=len(GRID_FLAGS)))
    for idx in rng.choice(len(combos), size=96, replace=False):
        cfg = spec.make_config(**dict(GRID_BASE, **dict(zip(GRID_FLAGS, combos[idx]))))
        w = synth.named_weights(spec.param_spec(cfg), seed=1)
        if g["ok"][idx]:
            out = mpl_oracle.forward(w, cfg, batch["poses"], batch["rays"], batch["centers"])
            assert out.shape == (2, 17, 3) and np.isfinite(out).all()
        else:
            with pytest.raises((RuntimeError, IndexError)):
                mpl_oracle.forward(w, cfg, batch["poses"], batch["rays"], batch["centers"])


def test_metric_matches_reference_semantics():
    rng = np.random.default_rng(1)
    pred, gt = rng.normal(size=(50, 17, 3)), rng.normal(size=(50, 17, 3))
    r = mpl_oracle.evaluate(pred, gt, output_in_meter=True, relative=False)
    d = np.sqrt((((pred - gt) * 100) ** 2).sum(-1))
    np.testing.assert_allclose(r["pjpe"], d.mean(0))
    np.testing.assert_allclose(r["mpjpe"], d.mean())
    rr = mpl_oracle.evaluate(pred, gt, output_in_meter=True, relative=True)
    pr, gr = pred - pred[:, :1], gt - gt[:, :1]
    np.testing.assert_allclose(rr["mpjpe"], np.sqrt((((pr - gr) * 100) ** 2).sum(-1)).mean())
    conf = np.ones((50, 17, 3))
    conf[3, 5] = 0
    rm = mpl_oracle.evaluate(pred, gt, conf_3d=conf)
    d2 = d.copy()
    d2[3, 5] = 0                                       # nansum -> 0 inside the joint norm, still counted in the mean
    np.testing.assert_allclose(rm["pjpe"], d2.mean(0))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_oracle_vs_live_reference():
    import torch
    m = ref_loader.load_model_module()
    for name in ("flag_hm0flags_small", "flag_confattn_multi", "flag_kadkhod", "cmu0_v2_d2"):
        case = CASES[name]
        cfg, weights, batch = make_inputs(case)
        model = m.MultiView_MPL(**case["kw"]).eval().double()
        model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
        V = cfg.V
        with torch.no_grad():
            ref = model([torch.from_numpy(batch["poses"][:, v]).double() for v in range(V)],
                        rays=[torch.from_numpy(batch["rays"][:, v]).double() for v in range(V)],
                        centers=[torch.from_numpy(batch["centers"][:, v]).double() for v in range(V)])
        ref = ref[0] if isinstance(ref, tuple) else ref
        out = mpl_oracle.forward(weights, cfg, batch["poses"], batch["rays"], batch["centers"])
        out = out[0] if isinstance(out, tuple) else out
        assert np.abs(out - ref.numpy()).max() < 1e-12


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_metric_vs_live_reference():
    ev = ref_loader.load_evaluate_module()
    rng = np.random.default_rng(2)
    pred, gt = rng.normal(size=(40, 17, 3)), rng.normal(size=(40, 17, 3))
    for mode in ("absolute", "relative"):
        a = ev.calc_mpjpe(pred, gt, mode=mode)
        b = mpl_oracle.calc_mpjpe(pred, gt, mode=mode)
        np.testing.assert_allclose(a[0], b[0])
        np.testing.assert_allclose(a[1], b[1])
    np.testing.assert_allclose(ev.calc_distance_per_dim(pred, gt)[0], mpl_oracle.calc_distance_per_dim(pred, gt)[0])


PROCRUSTES_MODES = [(True, "best"), (False, "best"), (True, False), (True, True)]


@pytest.mark.parametrize("scaling,reflection", PROCRUSTES_MODES)
def test_procrustes_restatement_matches_reference_golden(scaling, reflection):
    """tests/golden/procrustes.npz holds PoseUtils.procrustes outputs of the unmodified reference (oracle/make_goldens.py)."""
    g = np.load(os.path.join(GOLDEN, "procrustes.npz"))
    tag = f"s{int(scaling)}_r{reflection}"
    for i, (A, B) in enumerate(zip(g["A"], g["B"])):
        d, Z, tf = mpl_oracle.procrustes(A, B, scaling, reflection)
        assert abs(d - g["d_" + tag][i]) < 1e-12
        np.testing.assert_allclose(Z, g["Z_" + tag][i], atol=1e-11)
        np.testing.assert_allclose(tf["rotation"], g["R_" + tag][i], atol=1e-12)
        np.testing.assert_allclose(tf["scale"], g["scale_" + tag][i], rtol=1e-12)
        np.testing.assert_allclose(tf["translation"], g["t_" + tag][i], atol=1e-11)
        if reflection != "best":
            assert (np.linalg.det(tf["rotation"]) < 0) == bool(reflection)


def test_pmpjpe_sums_properties():
    """A similarity-transformed ground truth aligns back exactly; the aligned error never exceeds the raw one in rms."""
    g = np.load(os.path.join(GOLDEN, "procrustes.npz"))
    A = g["A"].astype(np.float64)
    rng = np.random.default_rng(3)
    Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    acc = mpl_oracle.pmpjpe_sums(1.7 * A @ Q + 0.3, A)
    J = A.shape[1]
    assert acc[J + 2] == len(A) and np.all(acc[:J] < 1e-9) and abs(acc[J]) < 1e-12
    np.testing.assert_allclose(acc[J + 1] / len(A), 1 / 1.7, rtol=1e-12)
    B = g["B"].astype(np.float64)
    for a, b in zip(A, B):
        _, Z, _ = mpl_oracle.procrustes(a, b)
        assert ((Z - a) ** 2).sum() <= ((b - a) ** 2).sum() + 1e-12


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_procrustes_vs_live_reference():
    u = ref_loader.load_pose_utils_module().PoseUtils()
    rng = np.random.default_rng(9)
    for i in range(20):
        A, B = rng.normal(size=(17, 3)), rng.normal(size=(17, 3))
        for scaling, reflection in PROCRUSTES_MODES:
            d, Z, tf = u.procrustes(A.copy(), B.copy(), scaling=scaling, reflection=reflection)
            d2, Z2, tf2 = mpl_oracle.procrustes(A, B, scaling, reflection)
            assert abs(d - d2) < 1e-13
            np.testing.assert_allclose(Z, Z2, atol=1e-12)
            np.testing.assert_allclose(tf["rotation"], tf2["rotation"], atol=1e-13)


@pytest.mark.parametrize("name", [n for n in CASES if not (CASES[n]["kw"].get("deep_head") or CASES[n]["kw"].get("head_kadkhod")
                                                            or CASES[n]["kw"].get("linear_weighted_mean"))])
def test_torch_cpu_port_matches_golden(name):
    """bench.py's CPU baseline (oracle/torch_port.py) is the reference's algorithm: fp32 round-off of the goldens."""
    import torch
    from oracle import torch_port
    cfg, weights, batch = make_inputs(CASES[name])
    g = load_golden(name)
    p = {k: torch.from_numpy(v) for k, v in weights.items()}
    out = torch_port.forward(p, cfg, *(torch.from_numpy(batch[k]) for k in ("poses", "rays", "centers"))).numpy()
    assert np.abs(out - g["out64_0"]).max() <= 2e-5 * max(np.abs(g["out64_0"]).max(), 1.0)
    assert np.abs(out - g["out32_0"]).max() <= 2e-5 * max(np.abs(g["out64_0"]).max(), 1.0)


def test_procrustes_invariances_random():
    """Properties PoseUtils.procrustes guarantees for any non-degenerate pair: orthogonal rotation, 0 <= d <= 1 with scaling,
    the aligned error is invariant under a similarity transform of the prediction, Z = scale * B R + t."""
    rng = np.random.default_rng(21)
    for _ in range(40):
        J = int(rng.integers(4, 25))
        A = rng.normal(size=(J, 3)) + rng.uniform(-3, 3, size=3)
        B = rng.normal(size=(J, 3))
        d, Z, tf = mpl_oracle.procrustes(A, B)
        R = tf["rotation"]
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
        assert -1e-12 <= d <= 1 + 1e-12
        np.testing.assert_allclose(Z, tf["scale"] * B @ R + tf["translation"], atol=1e-10)
        Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        d2, Z2, _ = mpl_oracle.procrustes(A, 2.5 * B @ Q - 4.0)
        assert abs(d - d2) < 1e-10
        np.testing.assert_allclose(Z, Z2, atol=1e-9)
        dr, Zr, tr = mpl_oracle.procrustes(A, B, scaling=False)          # rigid: scale reported as 1, residual larger
        assert tr["scale"] == 1 and ((Zr - A) ** 2).sum() >= ((Z - A) ** 2).sum() - 1e-10


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted (GPU box)")
@pytest.mark.parametrize("equal", [True, False])
def test_room_unscaling_is_the_reference_validate_tail(equal):
    """`oracle.room_unscale` against the reference's own statements: lines 474-488 of MPL/lib/core/function_mpl.py are read
    from the checkout and executed verbatim on a synthetic batch (the module itself cannot be imported: h5py / wandb)."""
    import textwrap
    import torch
    src = open(os.path.join(ref_loader.REF_ROOT, "MPL/lib/core/function_mpl.py")).read().splitlines()
    snippet = textwrap.dedent("\n".join(src[473:488]))
    assert snippet.startswith("preds = output.clone().cpu().numpy()") and "room_y_scale" in snippet
    rng = np.random.default_rng(3)
    out = rng.normal(size=(9, 17, 3)).astype(np.float32)
    tgt = rng.normal(size=(9, 17, 3)).astype(np.float32)
    meta0 = {"room_scaled": torch.ones(9), "room_x_scale": torch.full((9,), 3.25, dtype=torch.float64),
             "room_y_scale": torch.full((9,), 1.75, dtype=torch.float64)}
    room = {"room_x_scale": 3.25, "room_y_scale": 1.75}
    if equal:
        meta0["room_scaled_equal"] = torch.ones(9)
        meta0["room_center"] = torch.tensor([[0.5, -0.25, 0.9]] * 9, dtype=torch.float32)
        room = {"room_x_scale": 3.25, "room_center": [0.5, -0.25, 0.9]}
    env = {"output": torch.from_numpy(out), "t": torch.from_numpy(tgt), "meta": [meta0]}
    exec(snippet, env)
    p, g = mpl_oracle.room_unscale(out, tgt, room)
    np.testing.assert_array_equal(p, env["preds"])
    np.testing.assert_array_equal(g, env["gts"])
