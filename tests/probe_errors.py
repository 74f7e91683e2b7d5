"""Checker utility (test infrastructure, not collected by pytest): error of every tensor-core mode against the goldens.
    python tests/probe_errors.py [tf32 bf16 ...]     (PROBE_CASES=name,name to pick cases)"""
import sys, os
_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE)); sys.path.insert(0, _HERE)
import numpy as np
from conftest import load_golden
from gpu_util import build_module, run_module, rel_err, mpjpe_mm
from oracle.cases import CASES, make_inputs
NAMES = os.environ.get("PROBE_CASES")
for name in NAMES.split(",") if NAMES else ["hm0_v4_d12", "chosen_v4_d12", "cmu0_v2_d2", "cmu_v5_d2_hm0flags", "cmu_v5_d2_chosen", "sweep_viewtok_v2", "sweep_viewtok_v8", "kptok_v4_d12"]:
    case = CASES[name]; g = load_golden(name)
    cfg, weights, batch = make_inputs(case)
    t = batch["target"].astype(np.float64)
    for prec in (sys.argv[1:] or ["tf32"]):
        m = build_module(case["kw"], weights, prec)
        o = run_module(m, batch)[0]
        print(name, prec, "err %.2e" % rel_err(o, g["out64_0"]), "dmpjpe_mm %.3f" % abs(mpjpe_mm(o, t) - mpjpe_mm(g["out64_0"], t)), flush=True)
