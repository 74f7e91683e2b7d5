"""ORACLE — the parity case list shared by the golden generator and the tests (test infrastructure only).

Each case = constructor kwargs + rig kind + batch + seeds. Full-size architectures use tiny batches so the
fixtures stay small; flag-coverage cases use small widths (d=16, H=4, depth 2) so they run in milliseconds.
"""
from __future__ import annotations

from collections import OrderedDict

from openmpl_b200.spec import HM0_FLAGS, CHOSEN_FLAGS

_COMMON = dict(num_joints=17, embed_dim_ratio=32, num_heads=8, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1)
_SMALL = dict(num_joints=17, embed_dim_ratio=16, num_heads=4, depth=2, num_views=3, drop_path_rate=0.1)


def _case(kw, rig="h36m", batch=4, wseed=0, iseed=1):
    return dict(kw=kw, rig=rig, batch=batch, wseed=wseed, iseed=iseed)


CASES: "OrderedDict[str, dict]" = OrderedDict()
# --- shipped architectures (BASELINE.json configs 1-3) -----------------------------------
CASES["hm0_v4_d12"] = _case(dict(_COMMON, depth=12, num_views=4, **HM0_FLAGS), batch=4)
CASES["chosen_v4_d12"] = _case(dict(_COMMON, depth=12, num_views=4, **CHOSEN_FLAGS), batch=4)
CASES["cmu0_v2_d2"] = _case(dict(_COMMON, depth=2, num_views=2, **HM0_FLAGS), rig="cmu", batch=8)
CASES["cmu_v5_d2_hm0flags"] = _case(dict(_COMMON, depth=2, num_views=5, **HM0_FLAGS), rig="cmu", batch=8)
CASES["cmu_v5_d2_chosen"] = _case(dict(_COMMON, depth=2, num_views=5, **CHOSEN_FLAGS), rig="cmu", batch=8)
# --- view-count sweep (config 4): view tokens and view x keypoint tokens -------------------------------
for _v in (2, 3, 5, 6, 7, 8):
    CASES[f"sweep_viewtok_v{_v}"] = _case(dict(_COMMON, depth=2, num_views=_v, **HM0_FLAGS), batch=5, iseed=10 + _v)
    CASES[f"sweep_kptok_v{_v}"] = _case(dict(_COMMON, depth=2, num_views=_v, pose_3d_emb_learnable=True,
                                             FPT_blocks_view_keypoint_tokens=True), batch=5, iseed=20 + _v)
CASES["kptok_v4_d12"] = _case(dict(_COMMON, depth=12, num_views=4, pose_3d_emb_learnable=True,
                                   confidence_input_as_third=True, FPT_blocks_view_keypoint_tokens=True), batch=4)
# --- flag coverage at small width --------------------------------------------------
_FLAG_CASES = OrderedDict([
    ("plain", {}),
    ("learn3d", dict(pose_3d_emb_learnable=True)),
    ("conf3rd", dict(confidence_input_as_third=True)),
    ("addconf", dict(add_confidence_input=True)),
    ("multconf", dict(mult_confidence_emb=True)),
    ("addmultconf_multi", dict(add_confidence_input=True, mult_confidence_emb=True, multiple_spatial_blocks=True)),
    ("concatconf_q3", dict(concat_confidence_emb=True, add_confidence_input=True)),
    ("confattn", dict(confidence_as_attention_uncertainty_weight=True)),
    ("confattn_multi", dict(confidence_as_attention_uncertainty_weight=True, multiple_spatial_blocks=True,
                            confidence_input_as_third=True)),
    ("pos3d_spatial", dict(add_3D_pos_encoding_in_Spatial=True)),
    ("pos3d_spatial_learn", dict(add_3D_pos_encoding_in_Spatial=True, pose_3d_emb_learnable=True)),
    ("pos3d_spatial_rays", dict(add_3D_pos_encoding_in_Spatial=True, input_rays_as_token=True,
                                add_3D_pos_encoding_to_rays=True)),
    ("raytok_append", dict(input_rays_as_token=True)),
    ("raytok_append_learn", dict(input_rays_as_token=True, pose_3d_emb_learnable=True)),
    ("raytok_interleave_lin", dict(input_rays_as_token=True, add_3D_pos_encoding_to_rays=True)),
    ("linmean", dict(linear_weighted_mean=True, pose_3d_emb_learnable=True)),
    ("nospt", dict(no_transformer_spt=True)),
    ("nofpt", dict(no_transformer_fpt=True, input_rays_as_token=True)),
    ("nospt_nofpt", dict(no_transformer_spt=True, no_transformer_fpt=True)),
    ("conffpt", dict(confidence_in_FPT=True)),
    ("deephead", dict(deep_head=True, hidden_dim=64)),
    ("kadkhod", dict(head_kadkhod=True, hidden_dim=48)),
    ("kptok", dict(FPT_blocks_view_keypoint_tokens=True)),
    ("kptok_nofpt_rays", dict(FPT_blocks_view_keypoint_tokens=True, no_transformer_fpt=True,
                              input_rays_as_token=True)),
    ("noqkvbias_scale", dict(qkv_bias=False, qk_scale=0.3)),
    ("depth1", dict(depth=1)),
    ("mlp3", dict(mlp_ratio=3.0)),
    ("j13", dict(num_joints=13, pose_3d_emb_learnable=True)),
    ("hm0flags_small", dict(HM0_FLAGS)),
])
for _n, _f in _FLAG_CASES.items():
    CASES["flag_" + _n] = _case(dict(_SMALL, **_f), batch=5, wseed=3, iseed=7)

# the 12 shape-relevant boolean flags swept for the validity grid
GRID_FLAGS = ["confidence_input_as_third", "add_confidence_input", "pose_3d_emb_learnable", "linear_weighted_mean",
              "add_3D_pos_encoding_in_Spatial", "input_rays_as_token", "add_3D_pos_encoding_to_rays",
              "confidence_as_attention_uncertainty_weight", "multiple_spatial_blocks", "no_transformer_spt",
              "no_transformer_fpt", "FPT_blocks_view_keypoint_tokens"]
GRID_BASE = dict(num_joints=17, embed_dim_ratio=8, depth=1, num_heads=2, num_views=3)


def make_inputs(case):
    """(cfg, named weights, input batch dict) for a case — deterministic, numpy only."""
    from openmpl_b200 import spec, synth
    cfg = spec.make_config(**case["kw"])
    weights = synth.named_weights(spec.param_spec(cfg), seed=case["wseed"])
    rig = synth.make_rig(cfg.V, case["rig"])
    batch = synth.make_batch(case["batch"], rig, seed=case["iseed"])
    if cfg.J != synth.NUM_JOINTS:
        batch = {k: (v[:, :, :cfg.J] if k in ("poses", "rays") else v[:, :cfg.J] if k == "target" else v)
                 for k, v in batch.items()}
        batch = {k: v.copy() for k, v in batch.items()}
    return cfg, weights, batch
