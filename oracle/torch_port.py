"""ORACLE — test infrastructure, not product code.

The reference forward restated with the same CPU library calls the reference makes (torch F.linear / layer_norm /
softmax / gelu / matmul on CPU tensors, fp32), so that timing it on the GPU box's host cores is a fair stand-in for
"the reference's own CPU path" — the reference itself is pure Python/PyTorch and /root/reference does not exist on
the GPU box.  Functional (weights passed as a dict keyed by the reference's state_dict names), line-for-line the same
control flow as `oracle/mpl_oracle.py`, which cites `MPL/lib/models/multiview_mpl.py` per statement; default-head
configurations only (the CPU baseline is quoted on the shipped architectures).  Validated against the goldens in
tests/test_oracle_golden.py.  Only bench.py's cpu_baseline / --impl reference legs and tests/ import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _lin(x, p, pre):
    return F.linear(x, p[pre + "weight"], p.get(pre + "bias"))


def _attention(x, p, pre, H, qk_scale, conf_w):
    B, N, C = x.shape                                                     # multiview_mpl.py:53-67
    hd = C // H
    scale = qk_scale or hd ** -0.5
    qkv = _lin(x, p, pre + "qkv.").reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    if conf_w is not None:
        attn = attn * conf_w.unsqueeze(1)
    return _lin((attn @ v).transpose(1, 2).reshape(B, N, C), p, pre + "proj.")


def _block(x, p, pre, H, qk_scale=None, conf_w=None):
    C = x.shape[-1]                                                       # multiview_mpl.py:84-92
    x = x + _attention(F.layer_norm(x, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-6), p, pre + "attn.", H,
                       qk_scale, conf_w)
    h = F.gelu(_lin(F.layer_norm(x, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-6), p, pre + "mlp.fc1."))
    return x + _lin(h, p, pre + "mlp.fc2.")


def forward(p, cfg, poses, rays, centers):
    """poses/rays [B,V,J,3], centers [B,V,1,3] torch CPU tensors; p: name -> tensor. Returns [B,J,3]."""
    kw = cfg.kw
    if kw["deep_head"] or kw["head_kadkhod"] or kw["linear_weighted_mean"]:
        raise NotImplementedError("torch_port covers the default head only")
    b, d, J, V = poses.shape[0], cfg.d, cfg.J, cfg.V
    multi = kw["multiple_spatial_blocks"]
    xs = []
    with torch.no_grad():
        for i in range(V):                                                # multiview_mpl.py:458
            pose, ray, cen = poses[:, i], rays[:, i], centers[:, i]
            vs = f"{i}." if multi else ""
            conf_w = pose[:, :, 2:3].clone() if kw["confidence_as_attention_uncertainty_weight"] else None
            x = _lin(pose[:, :, 0:cfg.in_ch], p, f"Spatial_patch_to_embedding.{vs}")
            if kw["_add_conf"]:
                x = x + _lin(pose[:, :, 2:3], p, f"confidence_to_embedding.{vs}")
            if kw["_mult_conf"]:
                x = x * _lin(pose[:, :, 2:3], p, f"confidence_to_embedding.{vs}")
            x = x + p[f"Spatial_pos_embed.{i}" if multi else "Spatial_pos_embed"]
            if kw["add_3D_pos_encoding_in_Spatial"]:
                x = x + (p["pos_3d_embed"] if kw["pose_3d_emb_learnable"]
                         else _lin(F.normalize(ray - cen, dim=2, p=2), p, "pos_3d_linear."))
            if not kw["no_transformer_spt"]:
                for ix in range(cfg.depth):                               # :405-410 (last block twice)
                    pre = f"Spatial_blocks.{vs}{ix}."
                    if conf_w is not None:
                        x = _block(x, p, pre, cfg.H, kw["qk_scale"], conf_w)
                    if ix == cfg.depth - 1:
                        x = _block(x, p, pre, cfg.H, kw["qk_scale"])
                    x = _block(x, p, pre, cfg.H, kw["qk_scale"])
            x = F.layer_norm(x, (d,), p["Spatial_norm.weight"], p["Spatial_norm.bias"], 1e-6)
            if kw["confidence_in_FPT"]:
                x = x + _lin(pose[:, :, 2:3], p, "confidence_to_embedding_FPT.")
            if kw["add_3D_pos_encoding_to_rays"] and kw["input_rays_as_token"]:
                x = torch.cat([x, _lin(ray - cen, p, "ray_to_embedding.")], dim=2)
            if not kw["add_3D_pos_encoding_in_Spatial"]:
                pos = p["pos_3d_embed"] if kw["pose_3d_emb_learnable"] else _lin(F.normalize(ray - cen, dim=2, p=2), p, "pos_3d_linear.")
            else:
                pos = p["pos_3d_view_coding"]
            x = x + pos
            if not kw["add_3D_pos_encoding_to_rays"] and kw["input_rays_as_token"]:
                x = torch.cat([x, _lin(ray - cen, p, "ray_to_embedding.")], dim=1)
            xs.append(x.reshape(b, -1))
        x = torch.cat(xs, 1).reshape(b, cfg.fpt_tokens, -1)
        if not kw["no_transformer_fpt"]:
            for ix in range(cfg.depth):                                   # :420-423
                pre = f"blocks.{ix}."
                if ix == cfg.depth - 1:
                    x = _block(x, p, pre, cfg.H, kw["qk_scale"])
                x = _block(x, p, pre, cfg.H, kw["qk_scale"])
        if kw["input_rays_as_token"] and not kw["add_3D_pos_encoding_to_rays"]:
            x = x.reshape(b, V, 2, J, d)[:, :, 0].reshape(b, V, -1)
        elif kw["add_3D_pos_encoding_to_rays"]:
            x = x.reshape(b, V, J, 2 * d)[:, :, :, :d].reshape(b, V, -1)
        if kw["FPT_blocks_view_keypoint_tokens"]:
            x = x.reshape(b, V, -1)
        x = F.layer_norm(x, (cfg.E,), p["View_norm.weight"], p["View_norm.bias"], 1e-6)
        x = F.conv1d(x, p["weighted_mean.weight"], p["weighted_mean.bias"])                     # :445
        x = _lin(F.layer_norm(x, (cfg.E,), p["head.0.weight"], p["head.0.bias"], 1e-5), p, "head.1.")
        return x.reshape(b, -1, 3)
