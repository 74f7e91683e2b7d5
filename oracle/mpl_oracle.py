"""ORACLE — test infrastructure, not product code.

CPU restatement (numpy, fp64 by default) of the reference MPL lifter forward,
`/root/reference/MPL/lib/models/multiview_mpl.py`, and of its MPJPE metric,
`MPL/lib/core/evaluate.py:91-125` + `MPL/lib/core/function_mpl.py:670-687`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this package; the product path (`openmpl_b200/`) never does and fails loudly without its CUDA library.

Pinning: the reference ships no golden vectors or tests for this path, so the pins are
minted from the reference itself: `oracle/make_goldens.py` imports the unmodified reference module in the
build container and stores (inputs, fp32 + fp64 outputs) under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against every stored case (and against the live
reference whenever `/root/reference` is mounted).

Every function cites the reference lines it restates. Parameters are passed as a dict
name -> ndarray using the reference's `state_dict` names without the `features.` prefix.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf


def linear(x, p, prefix):
    """nn.Linear: x @ W^T + b."""
    y = x @ p[prefix + "weight"].T
    b = p.get(prefix + "bias")
    return y if b is None else y + b


def layer_norm(x, w, b, eps):
    """nn.LayerNorm over the last dim (biased variance)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def gelu(x):
    """nn.GELU() default = exact erf form (multiview_mpl.py:22,27)."""
    return 0.5 * x * (1.0 + erf(x * (1.0 / math.sqrt(2.0))))


def softmax(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


def attention(x, p, prefix, num_heads, qk_scale=None, conf_weights=None):
    """Attention.forward, multiview_mpl.py:53-67."""
    B, N, C = x.shape
    hd = C // num_heads
    scale = qk_scale or hd ** -0.5                                          # :46
    qkv = linear(x, p, prefix + "qkv.").reshape(B, N, 3, num_heads, hd).transpose(2, 0, 3, 1, 4)   # :55
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = softmax((q @ k.transpose(0, 1, 3, 2)) * scale)                  # :58-59
    if conf_weights is not None:
        attn = attn * conf_weights[:, None]                                 # :61-62  [B,1,N,1]: scales query rows
    y = (attn @ v).transpose(0, 2, 1, 3).reshape(B, N, C)                  # :64
    return linear(y, p, prefix + "proj.")                                   # :65


def block(x, p, prefix, num_heads, qk_scale=None, conf_weights=None):
    """Block.forward, multiview_mpl.py:84-92 (DropPath = identity in eval); block LayerNorms use eps 1e-6 (:139)."""
    h = layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-6)
    x = x + attention(h, p, prefix + "attn.", num_heads, qk_scale, conf_weights)
    h = layer_norm(x, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-6)
    h = gelu(linear(h, p, prefix + "mlp.fc1."))                             # Mlp.forward :31-37
    return x + linear(h, p, prefix + "mlp.fc2.")


def l2_normalize(x, eps=1e-12):
    """F.normalize(x, dim=-1, p=2): x / max(||x||, eps)."""
    n = np.sqrt((x * x).sum(axis=-1, keepdims=True))
    return x / np.maximum(n, eps)


def batch_norm_eval(x, p, prefix, eps=1e-5):
    """nn.BatchNorm1d in eval mode (running statistics)."""
    return (x - p[prefix + "running_mean"]) / np.sqrt(p[prefix + "running_var"] + eps) * p[prefix + "weight"] \
        + p[prefix + "bias"]


def spatial_forward_features(x, p, cfg, ray, center, view):
    """MultiView_MPL.Spatial_forward_features, multiview_mpl.py:349-414."""
    kw = cfg.kw
    multi = kw["multiple_spatial_blocks"]
    vs = f"{view}." if multi else ""
    conf_w = x[:, :, 2:3].copy() if kw["confidence_as_attention_uncertainty_weight"] else None   # :352-353
    emb = linear(x[:, :, 0:cfg.in_ch], p, f"Spatial_patch_to_embedding.{vs}")                   # :355-364
    if kw["_add_conf"]:
        emb = emb + linear(x[:, :, 2:3], p, f"confidence_to_embedding.{vs}")                     # :371-373
    if kw["_mult_conf"]:
        emb = emb * linear(x[:, :, 2:3], p, f"confidence_to_embedding.{vs}")                     # :374-376
    emb = emb + p[f"Spatial_pos_embed.{view}" if multi else "Spatial_pos_embed"]                 # :382-385
    x = emb
    if kw["add_3D_pos_encoding_in_Spatial"] and ray is not None and center is not None:         # :389-396
        if kw["pose_3d_emb_learnable"]:
            x = x + p["pos_3d_embed"]
        else:
            x = x + linear(l2_normalize(ray - center), p, "pos_3d_linear.")
    if not kw["no_transformer_spt"]:
        if multi and cfg.depth == 0:
            pass
        for ix in range(cfg.depth):                                                              # :405-410
            pre = f"Spatial_blocks.{vs}{ix}."
            if conf_w is not None:
                x = block(x, p, pre, cfg.H, kw["qk_scale"], conf_w)
            if ix == cfg.depth - 1:
                x = block(x, p, pre, cfg.H, kw["qk_scale"])                                      # Q1: last block twice
            x = block(x, p, pre, cfg.H, kw["qk_scale"])
    return layer_norm(x, p["Spatial_norm.weight"], p["Spatial_norm.bias"], 1e-6)                 # :412


def forward_features(x, p, cfg):
    """MultiView_MPL.forward_features, multiview_mpl.py:416-447."""
    kw = cfg.kw
    b = x.shape[0]
    if not kw["no_transformer_fpt"]:
        for ix in range(cfg.depth):                                                              # :420-423
            pre = f"blocks.{ix}."
            if ix == cfg.depth - 1:
                x = block(x, p, pre, cfg.H, kw["qk_scale"])
            x = block(x, p, pre, cfg.H, kw["qk_scale"])
    if kw["input_rays_as_token"] and not kw["add_3D_pos_encoding_to_rays"]:                      # :425-429
        x = x.reshape(b, cfg.V, 2, cfg.J, cfg.d)[:, :, 0].reshape(b, cfg.V, -1)
    elif kw["add_3D_pos_encoding_to_rays"]:                                                      # :430-434
        x = x.reshape(b, cfg.V, cfg.J, 2 * cfg.d)[:, :, :, :cfg.d].reshape(b, cfg.V, -1)
    if kw["FPT_blocks_view_keypoint_tokens"]:
        x = x.reshape(b, cfg.V, -1)                                                              # :436-437
    x = layer_norm(x, p["View_norm.weight"], p["View_norm.bias"], 1e-6)                          # :439
    if kw["linear_weighted_mean"]:
        x = linear(x.reshape(b, -1), p, "weighted_mean.")                                        # :441-443
    else:                                                                                        # Conv1d(V->1, k=1) :445
        x = np.einsum("v,bve->be", p["weighted_mean.weight"].reshape(-1), x) + p["weighted_mean.bias"]
    return x.reshape(b, 1, -1)


def forward(p, cfg, poses, rays, centers, dtype=np.float64):
    """MultiView_MPL.forward, multiview_mpl.py:450-525.

    poses/rays: [B,V,J,3]; centers: [B,V,1,3] (or lists of V per-view arrays). Returns [B,J,3]
    (or `(x, [x1, x2])` for head_kadkhod).
    """
    if cfg.error is not None:
        raise {"IndexError": IndexError}.get(cfg.error[0], RuntimeError)(cfg.error[1])
    kw = cfg.kw
    p = {k[len("features."):] if k.startswith("features.") else k: np.asarray(v, dtype=dtype)
         if np.asarray(v).dtype.kind == "f" else np.asarray(v) for k, v in p.items()}
    if isinstance(poses, (list, tuple)):
        poses, rays, centers = (np.stack([np.asarray(a) for a in t], axis=1) for t in (poses, rays, centers))
    poses, rays, centers = (np.asarray(a, dtype=dtype) for a in (poses, rays, centers))
    b = poses.shape[0]
    xs = []
    for i in range(poses.shape[1]):                                                              # :458
        pose, ray, cen = poses[:, i], rays[:, i], centers[:, i]
        x = spatial_forward_features(pose, p, cfg, ray, cen, i)                                  # :463
        if kw["confidence_in_FPT"]:
            x = x + linear(pose[:, :, 2:3], p, "confidence_to_embedding_FPT.")                  # :465-467
        if kw["add_3D_pos_encoding_to_rays"] and kw["input_rays_as_token"]:
            x = np.concatenate([x, linear(ray - cen, p, "ray_to_embedding.")], axis=2)           # :469-471
        if not kw["add_3D_pos_encoding_in_Spatial"]:                                             # :474-481
            if kw["pose_3d_emb_learnable"]:
                pos = p["pos_3d_embed"]
            else:
                pos = linear(l2_normalize(ray - cen), p, "pos_3d_linear.")
        else:
            pos = p["pos_3d_view_coding"]
        x = x + pos                                                                              # :483
        if not kw["add_3D_pos_encoding_to_rays"] and kw["input_rays_as_token"]:
            x = np.concatenate([x, linear(ray - cen, p, "ray_to_embedding.")], axis=1)           # :486-489
        xs.append(x.reshape(b, -1))
    xs = np.concatenate(xs, axis=1)
    xs = xs.reshape(b, cfg.fpt_tokens, -1)                                                       # :495-499
    x = forward_features(xs, p, cfg)                                                             # :505
    if kw["head_kadkhod"]:                                                                       # :506-516
        x = x.reshape(b, -1)

        def stage(z, s):
            pre = f"head.{s}."
            if s == 0:
                z = layer_norm(z, p[pre + "0.0.weight"], p[pre + "0.0.bias"], 1e-5)
                z = np.maximum(batch_norm_eval(linear(z, p, pre + "0.1."), p, pre + "0.2."), 0)
            else:
                z = np.maximum(batch_norm_eval(linear(z, p, pre + "0.0."), p, pre + "0.1."), 0)
            for k in (1, 2):
                z = np.maximum(batch_norm_eval(linear(z, p, pre + f"{k}.0."), p, pre + f"{k}.1."), 0)
            return linear(z, p, pre + "3.")
        x1 = stage(x, 0)
        x2 = stage(np.concatenate([x1, x], axis=1), 1)
        x3 = stage(np.concatenate([x2, x], axis=1), 2)
        return x3.reshape(b, -1, 3), [x1.reshape(b, -1, 3), x2.reshape(b, -1, 3)]
    if kw["deep_head"]:                                                                          # :517-519, :287-300
        z = layer_norm(x.reshape(b, -1), p["head.0.weight"], p["head.0.bias"], 1e-5)
        for lin, bn in (("head.1.", "head.2."), ("head.4.", "head.5."), ("head.7.", "head.8.")):
            z = np.maximum(batch_norm_eval(linear(z, p, lin), p, bn), 0)
        x = linear(z, p, "head.10.")
    else:                                                                                        # :283-286,:521  (Q2: eps 1e-5)
        x = linear(layer_norm(x, p["head.0.weight"], p["head.0.bias"], 1e-5), p, "head.1.")
    return x.reshape(b, -1, 3)                                                                   # :523


# --------------------------------------------------------------------------------------------------
# metric: MPL/lib/core/evaluate.py:91-125 and MPL/lib/core/function_mpl.py:670-687
# --------------------------------------------------------------------------------------------------

def calc_mpjpe(output, target, mode="absolute"):
    """evaluate.py:91-114 — per-joint mean L2 error [J] and its mean; NaNs are skipped inside the joint norm."""
    if mode == "relative":
        output = output - output[:, 0:1, :]
        target = target - target[:, 0:1, :]
    pjpe = np.sqrt(np.nansum((output - target) ** 2, axis=2)).mean(axis=0)
    return pjpe, pjpe.mean()


def calc_distance_per_dim(output, target):
    """evaluate.py:117-125."""
    distance = np.nanmean(np.abs(output - target), axis=0)
    return distance, distance.mean(axis=0)


def evaluate(pred, gt, output_in_meter=True, conf_3d=None, relative=False):
    """function_mpl.py:670-687: unit rule (x100 if OUTPUT_IN_METER), root-centring, conf masking, MPJPE."""
    pred = np.array(pred, dtype=np.float64)
    gt = np.array(gt, dtype=np.float64)
    if output_in_meter:
        pred, gt = pred * 100, gt * 100
    if relative:
        gt = gt - gt[:, 0:1, :]
        pred = pred - pred[:, 0:1, :]
    if conf_3d is not None:
        gt[conf_3d <= 0] = np.nan
        pred[conf_3d <= 0] = np.nan
    pjpe, mpjpe = calc_mpjpe(gt, pred, mode="relative" if relative else "absolute")
    dist_kp, dist = calc_distance_per_dim(pred, gt)
    return {"pjpe": pjpe, "mpjpe": mpjpe, "dist_per_dim_per_kp": dist_kp, "dist_per_dim": dist}


def room_unscale(preds, gts, room):
    """The un-scaling `validate()` applies to a batch before storing it (function_mpl.py:474-488), on fp32 arrays like
    there.  room: {'room_x_scale', 'room_center'} ('room_scaled_equal') or {'room_x_scale', 'room_y_scale'}."""
    preds = np.array(preds, dtype=np.float32)
    gts = np.array(gts, dtype=np.float32)
    if "room_center" in room:
        room_scale = float(room["room_x_scale"])
        room_center = np.asarray(room["room_center"], dtype=np.float32)
        preds = preds * room_scale + room_center
        gts = gts * room_scale + room_center
    else:
        preds[:, :, 0] = preds[:, :, 0] * float(room["room_x_scale"])
        preds[:, :, 1] = preds[:, :, 1] * float(room["room_y_scale"])
        gts[:, :, 0] = gts[:, :, 0] * float(room["room_x_scale"])
        gts[:, :, 1] = gts[:, :, 1] * float(room["room_y_scale"])
    return preds, gts


def metric_sums(pred, gt, conf_3d=None, output_in_meter=True):
    """The running sums `mpl_mpjpe_accumulate` keeps (layout in include/mpl_b200.h), restated from evaluate() above:
    finalising them must reproduce evaluate(relative=False/True) exactly."""
    pred = np.array(pred, dtype=np.float64)
    gt = np.array(gt, dtype=np.float64)
    B, J, _ = pred.shape
    if conf_3d is not None:
        conf_3d = np.broadcast_to(np.asarray(conf_3d).reshape(B, J, -1), (B, J, 3))
    a = evaluate(pred, gt, output_in_meter, conf_3d, relative=False)
    r = evaluate(pred, gt, output_in_meter, conf_3d, relative=True)
    cnt = np.full((J, 3), float(B)) if conf_3d is None else (conf_3d > 0).sum(axis=0).astype(np.float64)
    with np.errstate(invalid="ignore"):
        return np.concatenate([a["pjpe"] * B, r["pjpe"] * B, np.nan_to_num(a["dist_per_dim_per_kp"] * cnt).ravel(),
                               np.nan_to_num(r["dist_per_dim_per_kp"] * cnt).ravel(), cnt.ravel(), [float(B)]])


# --------------------------------------------------------------------------------------------------
# Procrustes-aligned error (P-MPJPE): MPL/lib/utils/pose_utils.py:61-143
# --------------------------------------------------------------------------------------------------

def procrustes(A, B, scaling=True, reflection="best"):
    """pose_utils.py:61-143 for equal column counts: similarity transform of B onto A.  Returns (d, Z, tform) with
    d the normalised residual, Z the transformed B and tform = {rotation, scale, translation} (Z = scale*B@R + t when
    scaling; with scaling=False the reference reports scale 1 and Z = ||B0|| * B0n @ R + mean(A))."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    a_bar, b_bar = A.mean(0), B.mean(0)                      # :89-93 remove translation
    A0, B0 = A - a_bar, B - b_bar
    ssx, ssy = (A0 ** 2).sum(), (B0 ** 2).sum()              # :96-101 remove scale
    a_norm, b_norm = np.sqrt(ssx), np.sqrt(ssy)
    A0, B0 = A0 / a_norm, B0 / b_norm
    U, s, Vt = np.linalg.svd(A0.T @ B0)                      # :107-110 optimum rotation of B
    V = Vt.T
    R = V @ U.T
    if reflection != "best":                                 # :112-120 force / forbid a reflection
        if bool(reflection) != (np.linalg.det(R) < 0):
            V = V.copy()
            s = s.copy()
            V[:, -1] *= -1
            s[-1] *= -1
            R = V @ U.T
    s_trace = s.sum()
    if scaling:                                              # :123-131
        scale = s_trace * a_norm / b_norm
        d = 1 - s_trace ** 2
        Z = a_norm * s_trace * (B0 @ R) + a_bar
    else:                                                    # :132-135
        scale = 1
        d = 1 + ssy / ssx - 2 * s_trace * b_norm / a_norm
        Z = b_norm * (B0 @ R) + a_bar
    return d, Z, {"rotation": R, "scale": scale, "translation": a_bar - scale * (b_bar @ R)}


def pmpjpe_sums(pred, gt, output_in_meter=True, scaling=True, reflection="best"):
    """The running sums `mpl_pmpjpe_accumulate` keeps (layout in include/mpl_b200.h): per pose, `procrustes(gt, pred)`
    after the unit rule of function_mpl.py:674-676, then the per-joint distances of calc_mpjpe (evaluate.py:104-110)
    between the aligned prediction and the ground truth."""
    pred = np.array(pred, dtype=np.float64)
    gt = np.array(gt, dtype=np.float64)
    if output_in_meter:
        pred, gt = pred * 100, gt * 100
    B, J, _ = pred.shape
    acc = np.zeros(J + 3)
    for b in range(B):
        d, Z, tf = procrustes(gt[b], pred[b], scaling, reflection)
        acc[:J] += np.sqrt(((Z - gt[b]) ** 2).sum(-1))
        acc[J] += d
        acc[J + 1] += tf["scale"] if scaling else 1.0
    acc[J + 2] = B
    return acc


# --------------------------------------------------------------------------------------------------
# input construction: MPL/lib/dataset/joints_dataset_mpl.py:615-648,701-715,762-772,817-820,872-904
# --------------------------------------------------------------------------------------------------

def normalize_screen_coordinates(X, w, h):
    """joints_dataset_mpl.py:817-820."""
    return (X / w) * 2 - np.array([1, h / w])


def build_inputs(pix, R, t, f, c, image_size):
    """Per-view model inputs from raw 2D detections, float64 like the dataset code, cast to float32 at the end.

    pix [B,V,J,3] (u, v, conf) pixels; R [V,3,3] world->cam; t [V,3] camera position (USE_T); f, c [V,2] pixels.
    Clip + confidence zeroing :709-715 (NO_AUGMENTATION branch), screen normalisation :762-764, intrinsics
    normalisation :615-623, rays R^T [(x-cx)/fx, (y-cy)/fy, 1] + t :872-898, centers = t^T :645-646.
    """
    pix = np.asarray(pix, dtype=np.float64)
    w, h = image_size
    B, V, J, _ = pix.shape
    poses = np.zeros((B, V, J, 3))
    rays = np.zeros((B, V, J, 3))
    centers = np.zeros((B, V, 1, 3))
    for v in range(V):
        joints = pix[:, v, :, :2].copy()
        vis = pix[:, v, :, 2].copy()
        vis = np.where(0 < joints[..., 0], vis, 0)
        vis = np.where(joints[..., 0] < w - 1, vis, 0)
        vis = np.where(0 < joints[..., 1], vis, 0)
        vis = np.where(joints[..., 1] < h - 1, vis, 0)
        joints[..., 0] = np.clip(joints[..., 0], 0, w - 1)
        joints[..., 1] = np.clip(joints[..., 1], 0, h - 1)
        cc = normalize_screen_coordinates(np.asarray(c[v], dtype=np.float64), w, h)
        ff = np.asarray(f[v], dtype=np.float64) / w * 2
        joints = normalize_screen_coordinates(joints, w, h)
        coords = joints.copy()
        coords[..., 0] = (coords[..., 0] - cc[0]) / ff[0]
        coords[..., 1] = (coords[..., 1] - cc[1]) / ff[1]
        cam = np.concatenate([coords, np.ones(coords.shape[:-1] + (1,))], axis=-1)
        world = cam @ np.asarray(R[v], dtype=np.float64) + np.asarray(t[v], dtype=np.float64)   # (R^T x)^T = x^T R
        poses[:, v] = np.concatenate([joints, vis[..., None]], axis=-1)
        rays[:, v] = world
        centers[:, v, 0] = t[v]
    return poses.astype(np.float32), rays.astype(np.float32), centers.astype(np.float32)
