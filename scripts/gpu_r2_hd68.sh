# usage: bash scripts/gpu_r2_hd68.sh TAG -- the fused QKV + attention kernel for 68-wide heads: unit + parity tests, then fused on / off benches
TAG=${1:-r2h68}
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_qkv_attn_gpu.py tests/test_parity_gpu.py -q -m gpu -x --timeout 120 2>&1 | tail -8
run() {  # name, env, bench args
  local name=$1 envs=$2; shift 2
  env $envs timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --parity-poses 64 "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "roof", round(d["roofline"]["frac"],3), "err", round(d["parity"]["max_abs_err_over_scale"],5))
    print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/${TAG}_$name.err").read()[-600:])
PY
}
run chosen_off MPL_QKV_ATTN_FUSION=0 --arch chosen
run chosen_on  MPL_QKV_ATTN_FUSION=1 --arch chosen
run chosen5_off MPL_QKV_ATTN_FUSION=0 --arch chosen --views 5 --depth 2
run chosen5_on  MPL_QKV_ATTN_FUSION=1 --arch chosen --views 5 --depth 2
run chosen_off2 MPL_QKV_ATTN_FUSION=0 --arch chosen
run chosen_on2  MPL_QKV_ATTN_FUSION=1 --arch chosen
