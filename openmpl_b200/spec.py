"""Constructor contract of the MPL lifter: kwargs -> derived dims, validity, parameter table.

Mirrors what `MultiView_MPL.__init__` (`MPL/lib/models/multiview_mpl.py:95-317`) derives from its 24+
keyword arguments, as data (a table of names/shapes) rather than as module-building code, so that the
host module, the oracle, the golden generator and the C-ABI all agree on one parameter list.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

# Keyword arguments of MultiView_MPL.__init__, in order, with the reference defaults (multiview_mpl.py:95-117).
CTOR_DEFAULTS = OrderedDict([
    ("num_joints", 17), ("in_chans", 2), ("embed_dim_ratio", 32), ("depth", 4), ("num_heads", 8),
    ("mlp_ratio", 2.0), ("qkv_bias", True), ("qk_scale", None), ("drop_rate", 0.0), ("attn_drop_rate", 0.0),
    ("drop_path_rate", 0.2), ("norm_layer", None), ("num_views", 5),
    ("add_confidence_input", False), ("mult_confidence_emb", False), ("concat_confidence_emb", False),
    ("confidence_input_as_third", False), ("pose_3d_emb_learnable", False), ("linear_weighted_mean", False),
    ("pos_embedding_type", "learnable"), ("add_3D_pos_encoding_in_Spatial", False),
    ("input_rays_as_token", False), ("add_3D_pos_encoding_to_rays", False),
    ("confidence_as_attention_uncertainty_weight", False), ("multiple_spatial_blocks", False),
    ("no_transformer_spt", False), ("no_transformer_fpt", False), ("confidence_in_FPT", False),
    ("deep_head", False), ("head_kadkhod", False), ("hidden_dim", 1024),
    ("FPT_blocks_view_keypoint_tokens", False),
])

BOOL_FLAGS = [k for k, v in CTOR_DEFAULTS.items() if isinstance(v, bool) and k != "qkv_bias"]

# "hm_0"/"cmu_0" yaml flag set and the "chosen" ablation.
HM0_FLAGS = dict(pose_3d_emb_learnable=True, confidence_input_as_third=True, input_rays_as_token=True,
                 multiple_spatial_blocks=True, add_3D_pos_encoding_to_rays=True)
CHOSEN_FLAGS = dict(pose_3d_emb_learnable=True)


@dataclass
class MplConfig:
    """Normalised constructor arguments plus the dims derived from them."""
    kw: dict
    J: int = 0
    V: int = 0
    d: int = 0
    H: int = 0
    depth: int = 0
    in_ch: int = 0                # columns of the pose consumed by the joint embedding (2 or 3)
    conf_emb: bool = False        # confidence_to_embedding exists (add or mult mode live, Q3 applied)
    tok_w: int = 0                # per-view flattened token width entering the FPT: J*d or 2*J*d
    fpt_dim: int = 0              # channel width of the FPT blocks
    fpt_tokens: int = 0           # tokens per pose in the FPT
    E: int = 0                    # J*d : width of View_norm / head input
    pos3d_lin_out: int = 0
    pos3d_w: int = 0
    spt_hidden: int = 0
    fpt_hidden: int = 0
    ray_layout: str = "none"      # "none" | "interleave" (cat dim=2, add_to_rays) | "append" (cat dim=1)
    error: tuple | None = None    # (exception class name, message) raised at first forward, as the reference does

    def flag(self, name):
        return self.kw[name]


def make_config(**kwargs) -> MplConfig:
    kw = OrderedDict(CTOR_DEFAULTS)
    for k, v in kwargs.items():
        if k not in kw:
            raise TypeError(f"MultiView_MPL.__init__() got an unexpected keyword argument {k!r}")
        kw[k] = v
    if kw["norm_layer"] is not None:
        raise NotImplementedError("norm_layer must be None (LayerNorm eps=1e-6, multiview_mpl.py:139)")
    c = MplConfig(kw=dict(kw))
    c.J, c.V, c.d, c.H, c.depth = kw["num_joints"], kw["num_views"], kw["embed_dim_ratio"], kw["num_heads"], kw["depth"]
    c.in_ch = kw["in_chans"] + 1 if kw["confidence_input_as_third"] else kw["in_chans"]
    # Q3: concat_confidence_emb switches all three confidence-embedding modes off (multiview_mpl.py:173-176)
    add_c, mult_c = kw["add_confidence_input"], kw["mult_confidence_emb"]
    if kw["concat_confidence_emb"]:
        add_c = mult_c = False
    c.kw["_add_conf"], c.kw["_mult_conf"] = add_c, mult_c
    c.conf_emb = add_c or mult_c
    rays_tok, to_rays = kw["input_rays_as_token"], kw["add_3D_pos_encoding_to_rays"]
    c.E = c.d * c.J
    c.tok_w = c.E * (2 if rays_tok else 1)
    c.ray_layout = "none" if not rays_tok else ("interleave" if to_rays else "append")
    if kw["FPT_blocks_view_keypoint_tokens"]:
        c.fpt_dim, c.fpt_tokens = c.d, c.V * c.J
    else:
        c.fpt_dim, c.fpt_tokens = c.tok_w, c.V
    c.pos3d_lin_out = 2 * c.d if (to_rays and not kw["add_3D_pos_encoding_in_Spatial"]) else c.d
    c.pos3d_w = 2 * c.d if to_rays else c.d
    c.spt_hidden = int(c.d * kw["mlp_ratio"])
    c.fpt_hidden = int(c.fpt_dim * kw["mlp_ratio"])
    c.error = _first_forward_error(c)
    return c


def _first_forward_error(c: MplConfig):
    """Flag combinations whose first forward raises in the reference."""
    kw = c.kw
    if kw["in_chans"] != 2:
        return ("RuntimeError", "in_chans must be 2: the forward slices pose[:, :, 0:2|0:3] (multiview_mpl.py:359-364)")
    # (with depth 0 the reference builds no blocks, so nothing reshapes: multiview_mpl.py:236-259)
    if c.d % c.H != 0 and not kw["no_transformer_spt"] and c.depth > 0:
        return ("RuntimeError", "embed_dim_ratio must be divisible by num_heads (reshape at multiview_mpl.py:55)")
    if kw["no_transformer_spt"] and kw["multiple_spatial_blocks"]:
        return ("IndexError", "index 0 is out of range (Spatial_blocks is empty, multiview_mpl.py:401)")
    if kw["add_3D_pos_encoding_to_rays"] and not kw["input_rays_as_token"]:
        return ("RuntimeError", f"The size of tensor a ({c.d}) must match the size of tensor b ({2 * c.d}) at "
                                "non-singleton dimension 2 (multiview_mpl.py:483)")
    if (kw["input_rays_as_token"] and kw["add_3D_pos_encoding_to_rays"] and kw["add_3D_pos_encoding_in_Spatial"]
            and kw["pose_3d_emb_learnable"]):
        return ("RuntimeError", f"The size of tensor a ({c.d}) must match the size of tensor b ({2 * c.d}) at "
                                "non-singleton dimension 2 (multiview_mpl.py:396)")
    if kw["input_rays_as_token"] and kw["FPT_blocks_view_keypoint_tokens"] and not kw["no_transformer_fpt"]:
        return ("RuntimeError", f"Given normalized_shape=[{c.d}], expected input with shape [*, {c.d}], but got "
                                f"input of width {2 * c.d} (multiview_mpl.py:75,497)")
    if not kw["no_transformer_fpt"] and c.fpt_dim % c.H != 0 and c.depth > 0:
        return ("RuntimeError", "FPT width must be divisible by num_heads (reshape at multiview_mpl.py:55)")
    return None


def _block(spec, prefix, dim, hidden, qkv_bias):
    spec[prefix + "norm1.weight"] = ((dim,), "norm_w", dim)
    spec[prefix + "norm1.bias"] = ((dim,), "norm_b", dim)
    spec[prefix + "attn.qkv.weight"] = ((3 * dim, dim), "linear_w", dim)
    if qkv_bias:
        spec[prefix + "attn.qkv.bias"] = ((3 * dim,), "linear_b", dim)
    spec[prefix + "attn.proj.weight"] = ((dim, dim), "linear_w", dim)
    spec[prefix + "attn.proj.bias"] = ((dim,), "linear_b", dim)
    spec[prefix + "norm2.weight"] = ((dim,), "norm_w", dim)
    spec[prefix + "norm2.bias"] = ((dim,), "norm_b", dim)
    spec[prefix + "mlp.fc1.weight"] = ((hidden, dim), "linear_w", dim)
    spec[prefix + "mlp.fc1.bias"] = ((hidden,), "linear_b", dim)
    spec[prefix + "mlp.fc2.weight"] = ((dim, hidden), "linear_w", hidden)
    spec[prefix + "mlp.fc2.bias"] = ((dim,), "linear_b", hidden)


def _linear(spec, prefix, out_f, in_f):
    spec[prefix + "weight"] = ((out_f, in_f), "linear_w", in_f)
    spec[prefix + "bias"] = ((out_f,), "linear_b", in_f)


def _bn(spec, prefix, n):
    spec[prefix + "weight"] = ((n,), "norm_w", n)
    spec[prefix + "bias"] = ((n,), "norm_b", n)
    spec[prefix + "running_mean"] = ((n,), "bn_mean", n)
    spec[prefix + "running_var"] = ((n,), "bn_var", n)
    spec[prefix + "num_batches_tracked"] = ((), "count", 1)


BUFFER_KINDS = ("bn_mean", "bn_var", "count")


def param_spec(c: MplConfig, prefix: str = "") -> "OrderedDict[str, tuple]":
    """name -> (shape, kind, fan_in) for every state_dict entry of `MultiView_MPL` (names as in the reference).

    Follows the registration order of multiview_mpl.py:158-317 (own parameters first, then children).
    """
    kw, d, J, V, E = c.kw, c.d, c.J, c.V, c.E
    multi = kw["multiple_spatial_blocks"]
    s: OrderedDict = OrderedDict()
    views = [f"{v}." for v in range(V)] if multi else [""]
    if not multi:
        s["Spatial_pos_embed"] = ((1, J, d), "pos", d)
    s["pos_3d_embed"] = ((1, J, c.pos3d_w), "pos", d)
    s["pos_3d_view_coding"] = ((1, J, c.pos3d_w), "pos", d)
    for v in views:
        _linear(s, f"Spatial_patch_to_embedding.{v}", d, c.in_ch)
    if c.conf_emb:
        for v in views:
            _linear(s, f"confidence_to_embedding.{v}", d, 1)
    if multi:
        for v in range(V):
            s[f"Spatial_pos_embed.{v}"] = ((1, J, d), "pos", d)
    _linear(s, "pos_3d_linear.", c.pos3d_lin_out, 3)
    if kw["input_rays_as_token"]:
        _linear(s, "ray_to_embedding.", d, 3)
    if kw["confidence_in_FPT"]:
        _linear(s, "confidence_to_embedding_FPT.", d, 1)
    if not kw["no_transformer_spt"]:
        for v in views:
            for l in range(c.depth):
                _block(s, f"Spatial_blocks.{v}{l}.", d, c.spt_hidden, kw["qkv_bias"])
    if not kw["no_transformer_fpt"]:
        for l in range(c.depth):
            _block(s, f"blocks.{l}.", c.fpt_dim, c.fpt_hidden, kw["qkv_bias"])
    s["Spatial_norm.weight"] = ((d,), "norm_w", d)
    s["Spatial_norm.bias"] = ((d,), "norm_b", d)
    s["View_norm.weight"] = ((E,), "norm_w", E)
    s["View_norm.bias"] = ((E,), "norm_b", E)
    if kw["linear_weighted_mean"]:
        _linear(s, "weighted_mean.", E, V * E)
    else:
        s["weighted_mean.weight"] = ((1, V, 1), "linear_w", V)
        s["weighted_mean.bias"] = ((1,), "linear_b", V)
    out_dim, Hd = 3 * J, kw["hidden_dim"]
    if kw["head_kadkhod"]:
        for stage in range(3):
            first_in = E if stage == 0 else out_dim + E
            p = f"head.{stage}."
            if stage == 0:
                s[p + "0.0.weight"] = ((E,), "norm_w", E)
                s[p + "0.0.bias"] = ((E,), "norm_b", E)
                _linear(s, p + "0.1.", Hd, first_in)
                _bn(s, p + "0.2.", Hd)
            else:
                _linear(s, p + "0.0.", Hd, first_in)
                _bn(s, p + "0.1.", Hd)
            for k in (1, 2):
                _linear(s, p + f"{k}.0.", Hd, Hd)
                _bn(s, p + f"{k}.1.", Hd)
            _linear(s, p + "3.", out_dim, Hd)
    elif kw["deep_head"]:
        s["head.0.weight"] = ((E,), "norm_w", E)
        s["head.0.bias"] = ((E,), "norm_b", E)
        _linear(s, "head.1.", Hd, E)
        _bn(s, "head.2.", Hd)
        _linear(s, "head.4.", Hd, Hd)
        _bn(s, "head.5.", Hd)
        _linear(s, "head.7.", Hd, Hd)
        _bn(s, "head.8.", Hd)
        _linear(s, "head.10.", out_dim, Hd)
    else:
        s["head.0.weight"] = ((E,), "norm_w", E)
        s["head.0.bias"] = ((E,), "norm_b", E)
        _linear(s, "head.1.", out_dim, E)
    if prefix:
        s = OrderedDict((prefix + k, v) for k, v in s.items())
    return s


def flops_per_pose(c: MplConfig) -> float:
    """Algorithmic FLOPs of one forward per pose, counting depth+1 block applications."""
    kw = c.kw

    def blk(n, dim, hidden):
        return 2 * n * dim * 3 * dim + 4 * n * n * dim + 2 * n * dim * dim + 4 * n * dim * hidden

    apps = c.depth + 1 if c.depth > 0 else 0
    total = 0.0
    if not kw["no_transformer_spt"]:
        spt_apps = apps + (c.depth if kw["confidence_as_attention_uncertainty_weight"] else 0)
        total += c.V * spt_apps * blk(c.J, c.d, c.spt_hidden)
    if not kw["no_transformer_fpt"]:
        total += apps * blk(c.fpt_tokens, c.fpt_dim, c.fpt_hidden)
    total += 2 * c.V * c.J * c.in_ch * c.d                      # joint embedding
    if kw["input_rays_as_token"]:
        total += 2 * c.V * c.J * 3 * c.d
    total += 2 * c.E * 3 * c.J                                   # head Linear (default head)
    return float(total)
