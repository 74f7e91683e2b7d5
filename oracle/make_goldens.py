"""ORACLE — mint the golden vectors from the UNMODIFIED reference (run in the build container only).

    python -m oracle.make_goldens            # writes tests/golden/*.npz and tests/golden/validity_grid.npz
    python -m oracle.make_goldens --only-procrustes     # just tests/golden/procrustes.npz (PoseUtils.procrustes)

For every case in `oracle/cases.py`: build the reference `MultiView_MPL(**kw)`, load the name-keyed deterministic
weights (`openmpl_b200.synth.named_weights` — regenerated, not stored), run the seeded synthetic inputs through
it in fp32 and in fp64, and store inputs + both outputs. The validity grid stores, for all 4096 combinations of the
12 shape-relevant flags, whether the reference's first forward succeeds.
"""
from __future__ import annotations

import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                               # noqa: E402
from oracle.cases import CASES, GRID_FLAGS, GRID_BASE, make_inputs   # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def run_reference(kw, weights, batch, dtype):
    m = ref_loader.load_model_module()
    model = m.MultiView_MPL(**kw).eval()
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()}
    model.load_state_dict(sd, strict=True)
    model = model.to(dtype)
    V = batch["poses"].shape[1]
    args = {k: [torch.from_numpy(batch[k][:, v]).to(dtype) for v in range(V)] for k in ("poses", "rays", "centers")}
    with torch.no_grad():
        out = model(args["poses"], rays=args["rays"], centers=args["centers"])
    if isinstance(out, tuple):
        return [out[0].numpy()] + [o.numpy() for o in out[1]]
    return [out.numpy()]


PROCRUSTES_MODES = [(True, "best"), (False, "best"), (True, False), (True, True)]   # (scaling, reflection)


def procrustes_inputs(n=48, J=17, seed=11):
    """Ground-truth poses A and predictions B: noisy similarity transforms of A (some mirrored), plus unrelated pairs."""
    rng = np.random.default_rng(seed)
    A = rng.normal(0.0, 0.3, size=(n, J, 3)) + rng.uniform(-2, 2, size=(n, 1, 3))
    B = np.empty_like(A)
    for i in range(n):
        Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        if i % 4 == 1:
            Q[:, 0] *= -1                                    # a mirrored prediction
        B[i] = rng.uniform(0.5, 2.0) * (A[i] @ Q) + rng.normal(size=3) + rng.normal(0, 0.02 * (1 + i % 5), size=(J, 3))
        if i % 6 == 5:
            B[i] = rng.normal(0.0, 0.3, size=(J, 3))         # unrelated pose
    return A.astype(np.float32), B.astype(np.float32)


def mint_procrustes():
    """PoseUtils.procrustes (pose_utils.py:61-143) of the unmodified reference on seeded pose pairs."""
    u = ref_loader.load_pose_utils_module().PoseUtils()
    A, B = procrustes_inputs()
    arrays = dict(A=A, B=B)
    for scaling, refl in PROCRUSTES_MODES:
        tag = f"s{int(scaling)}_r{refl}"
        res = [u.procrustes(a.astype(np.float64), b.astype(np.float64), scaling=scaling, reflection=refl)
               for a, b in zip(A, B)]
        arrays[f"d_{tag}"] = np.array([r[0] for r in res])
        arrays[f"Z_{tag}"] = np.stack([r[1] for r in res])
        arrays[f"R_{tag}"] = np.stack([r[2]["rotation"] for r in res])
        arrays[f"scale_{tag}"] = np.array([float(r[2]["scale"]) for r in res])
        arrays[f"t_{tag}"] = np.stack([r[2]["translation"] for r in res])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "procrustes.npz"), **arrays)
    print("procrustes:", A.shape, "modes", PROCRUSTES_MODES)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    mint_procrustes()
    if "--only-procrustes" in sys.argv:
        return
    torch.set_num_threads(os.cpu_count())
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--only=")]       # --only=name,name: just these cases
    for name, case in CASES.items():
        if only and name not in only[0]:
            continue
        cfg, weights, batch = make_inputs(case)
        out32 = run_reference(case["kw"], weights, batch, torch.float32)
        out64 = run_reference(case["kw"], weights, batch, torch.float64)
        arrays = dict(poses=batch["poses"], rays=batch["rays"], centers=batch["centers"], target=batch["target"])
        for i, (a, b) in enumerate(zip(out32, out64)):
            arrays[f"out32_{i}"] = a
            arrays[f"out64_{i}"] = b
        arrays["meta"] = np.array(json.dumps(dict(case=name, kw=case["kw"], rig=case["rig"], wseed=case["wseed"],
                                                  iseed=case["iseed"], torch=torch.__version__)))
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
        print(f"{name}: out {out64[0].shape} scale {np.abs(out64[0]).max():.3f} "
              f"fp32-vs-fp64 {np.abs(out32[0] - out64[0]).max():.2e}")
    if only:
        return
    # validity grid: does the reference's constructor + first forward succeed?
    m = ref_loader.load_model_module()
    ok = np.zeros(1 << len(GRID_FLAGS), dtype=np.uint8)
    err = {}
    V, J = GRID_BASE["num_views"], GRID_BASE["num_joints"]
    x = [torch.rand(2, J, 3) for _ in range(V)]
    r = [torch.rand(2, J, 3) for _ in range(V)]
    c = [torch.rand(2, 1, 3) for _ in range(V)]
    for idx, bits in enumerate(itertools.product((False, True), repeat=len(GRID_FLAGS))):
        kw = dict(GRID_BASE, **dict(zip(GRID_FLAGS, bits)))
        try:
            with torch.no_grad():
                m.MultiView_MPL(**kw).eval()(x, rays=r, centers=c)
            ok[idx] = 1
        except Exception as e:                               # noqa: BLE001
            err[type(e).__name__] = err.get(type(e).__name__, 0) + 1
    np.savez_compressed(os.path.join(GOLDEN_DIR, "validity_grid.npz"), ok=ok,
                        meta=np.array(json.dumps(dict(flags=GRID_FLAGS, base=GRID_BASE))))
    print("validity grid:", int(ok.sum()), "of", ok.size, "run;", err)


if __name__ == "__main__":
    main()
