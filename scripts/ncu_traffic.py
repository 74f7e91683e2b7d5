#!/usr/bin/env python
"""Regenerate profiles/ncu_traffic.json (what bench.py reports as roofline.traffic) from an `ncu --set full` capture of the
four FPT GEMM launches of one block, stamped with the build it was taken from.

    python scripts/ncu_traffic.py gpurun_out/<capture>.ncu-rep [arch] [profiles/<summary>.md]
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"4": "qkv_ln_bias", "5": "fc1_ln_bias_gelu", "6": "proj_residual_emit", "7": "fc2_residual_emit"}


def main():
    rep = sys.argv[1]
    arch = sys.argv[2] if len(sys.argv) > 2 else "hm0"
    src = sys.argv[3] if len(sys.argv) > 3 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, k):
        v, u = float(r[ix[k]].replace(",", "")), units[ix[k]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    per, tens = {}, {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        if "qkv_attn_kernel" in name:
            key = "qkv_attention_fused"
        elif "gemm_tcgen05_kernel" in name:
            epi = name.split("gemm_tcgen05_kernel<")[1].split(">")[0].replace("(int)", "").split(",")[2].strip()
            key = NAMES.get(epi, "epi" + epi)
        else:
            continue
        per[key] = (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) / 1e6
        tens[key] = float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
    lib = os.path.join(ROOT, "openmpl_b200", "libmpl_b200.so")
    rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    data[arch] = {"gemm_tcgen05_kernel": {
        "bytes_per_launch": 1e6 * sum(per.values()) / max(len(per), 1),
        "note": "mean over the FPT GEMM instantiations of one block (" + ", ".join(f"{k} {v:.0f} MB" for k, v in per.items())
                + f"), ncu --set full, chunk of 32768 poses, {os.path.basename(src)}",
        "per_instantiation_mb": per, "tensor_pipe_pct_ncu": tens,
        "capture": os.path.basename(rep), "git_rev_at_capture": rev,
        "lib_sha256_16": hashlib.sha256(open(lib, "rb").read()).hexdigest()[:16] if os.path.isfile(lib) else None}}
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(data[arch], indent=1))


if __name__ == "__main__":
    main()
