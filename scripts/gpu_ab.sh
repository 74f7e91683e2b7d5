# usage: bash scripts/gpu_ab.sh TAG "ENV1=.. ENV2=.." "ENV..." ...   -- bench (no tests) under several env settings (same box: A/B)
TAG=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  env $envs timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --parity-poses 64 > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$i.json").read().strip().splitlines()[-1])
    print("[$envs]", "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],2), "roof", round(d["roofline"]["frac"],3), "err", round(d["parity"]["max_abs_err_over_scale"],5))
    print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
except Exception as e:
    print("[$envs] ERR", e); print(open("gpurun_out/${TAG}_$i.err").read()[-600:])
PY
  i=$((i+1))
done
