mkdir -p gpurun_out
TAG=${1:-r2i}
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout -k 5 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --precision tf32 --parity-poses 64 > gpurun_out/${TAG}_bench_tf32.json 2> gpurun_out/${TAG}_bench_tf32.err; echo "bench tf32 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json","gpurun_out/${TAG}_bench_tf32.json"):
  try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(d["dtype"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "roof", round(d["roofline"]["frac"],3), "parity", d["parity"])
    print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
    print("   latency", d.get("latency_b256")); print("   fp32grade", d.get("fp32_grade_mode"))
  except Exception as e:
    print("ERR", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
