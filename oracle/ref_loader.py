"""ORACLE helper — import the UNMODIFIED reference model from /root/reference (build container only).

The reference needs `timm` and `easydict`, which are not installed; of the seven timm names it imports
(`multiview_mpl.py:13-16`) only `DropPath` is live (`:79`, identity in eval) and `trunc_normal_` sits in a dead
branch (`:599`). Tiny stand-in modules are registered in `sys.modules` so the file loads as-is
(SURVEY.md Appendix A). Nothing here is used on the GPU box: `/root/reference` does not exist there.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("MPL_REFERENCE_ROOT", "/root/reference")
MODEL_FILE = os.path.join(REF_ROOT, "MPL/lib/models/multiview_mpl.py")
# byte-for-byte copy of the model file made by oracle/stage_reference.py (git-ignored; travels to the GPU box)
STAGE_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
STAGED_MODEL_FILE = os.path.join(STAGE_ROOT, "MPL/lib/models/multiview_mpl.py")


def available() -> bool:
    """The whole reference checkout is mounted (build container)."""
    return os.path.isfile(MODEL_FILE)


def model_file() -> str | None:
    """The unmodified reference model file: from the checkout, else the staged copy (GPU box), else None."""
    for f in (MODEL_FILE, STAGED_MODEL_FILE):
        if os.path.isfile(f):
            return f
    return None


def _install_stubs():
    import torch

    if "timm" not in sys.modules:
        class DropPath(torch.nn.Module):                      # timm 0.6.13 semantics
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1 - self.drop_prob
                mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
                return x * mask / keep

        mods = {n: types.ModuleType(n) for n in
                ("timm", "timm.data", "timm.models", "timm.models.helpers", "timm.models.layers",
                 "timm.models.registry")}
        mods["timm.data"].IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
        mods["timm.data"].IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
        mods["timm.models.helpers"].load_pretrained = lambda *a, **k: None
        mods["timm.models.layers"].DropPath = DropPath
        mods["timm.models.layers"].to_2tuple = lambda x: (x, x)
        mods["timm.models.layers"].trunc_normal_ = torch.nn.init.trunc_normal_
        mods["timm.models.registry"].register_model = lambda f: f
        sys.modules.update(mods)
    if "easydict" not in sys.modules:
        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in dict(d or {}, **kw).items():
                    self[k] = v

            def __setitem__(self, k, v):
                if isinstance(v, dict) and not isinstance(v, EasyDict):
                    v = EasyDict(v)
                super().__setitem__(k, v)

            __setattr__ = __setitem__

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

        m = types.ModuleType("easydict")
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m


_cached = None


def load_model_module():
    """The reference `multiview_mpl` python module, loaded from its file unmodified."""
    global _cached
    if _cached is None:
        f = model_file()
        if f is None:
            raise FileNotFoundError(MODEL_FILE)
        _install_stubs()
        spec = importlib.util.spec_from_file_location("ref_multiview_mpl", f)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cached = mod
    return _cached


def load_config_module():
    """The reference `core.config` module (global EasyDict + update_config)."""
    _install_stubs()
    lib = os.path.join(REF_ROOT, "MPL/lib")
    spec = importlib.util.spec_from_file_location("ref_core_config", os.path.join(lib, "core/config.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_evaluate_module():
    """The reference `core.evaluate` module (calc_mpjpe)."""
    _install_stubs()
    lib = os.path.join(REF_ROOT, "MPL/lib")
    if lib not in sys.path:
        sys.path.insert(0, lib)
    spec = importlib.util.spec_from_file_location("ref_core_evaluate", os.path.join(lib, "core/evaluate.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_pose_utils_module():
    """The reference `utils.pose_utils` module (PoseUtils.procrustes); needs numpy + sklearn only."""
    import warnings
    spec = importlib.util.spec_from_file_location("ref_utils_pose_utils",
                                                  os.path.join(REF_ROOT, "MPL/lib/utils/pose_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)      # `reflection is not 'best'` in the reference source
        spec.loader.exec_module(mod)
    return mod
