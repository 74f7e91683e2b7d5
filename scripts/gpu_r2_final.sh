# final evidence of the round on one GPU: full -m gpu suite, smoke(), bench (ours + reference arm), launch list
mkdir -p gpurun_out
TAG=${1:-r2final}
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${TAG}_smoke.log
timeout -k 5 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"
timeout -k 5 500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --parity-poses 0 > gpurun_out/${TAG}_launches.log 2>&1; wc -l gpurun_out/${TAG}_launches.csv
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("bf16 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "roof", round(d["roofline"]["frac"],3), "parity", d["parity"], "clocks", d["clocks"])
print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
print("   latency", d.get("latency_b256")); print("   fp32grade", d.get("fp32_grade_mode")); print("   cpu", d.get("cpu_baseline")); print("   mem", d.get("memory_bound_kernels"))
r=json.loads(open("gpurun_out/${TAG}_ref.json").read().strip().splitlines()[-1]); print("reference arm", round(r["value"]), r["cpu_baseline"])
PY
