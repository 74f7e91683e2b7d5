mkdir -p gpurun_out
TAG=${1:-r2h}
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/${TAG}_tests.log
timeout -k 5 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("bf16 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "roof", round(d["roofline"]["frac"],3), "parity", d["parity"])
    print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
    print("   latency", d.get("latency_b256")); print("   fp32grade", d.get("fp32_grade_mode")); print("   cpu", d.get("cpu_baseline"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${TAG}_bench.err").read()[-2500:])
PY
timeout -k 5 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_ref.json
