# usage: bash scripts/gpu_tests_bench.sh [tag]  -- full -m gpu suite + default bench without the CPU leg
TAG=${1:-run}
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/${TAG}_rc.txt
timeout -k 5 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bench rc=$?" >> gpurun_out/${TAG}_rc.txt
cat gpurun_out/${TAG}_rc.txt; tail -8 gpurun_out/${TAG}_tests.log
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_bf16.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "roof", d["roofline"]["frac"], "parity", d["parity"])
    for k,v in d["breakdown"].items(): print("   ", k, round(v["ms_per_step"],3), v["launches_per_step"])
    print({k: round(v["frac_of_hbm_peak"],3) for k,v in d["memory_bound_kernels"].items()})
except Exception as e:
    print("ERR", e)
PY
tail -3 gpurun_out/${TAG}_bench_bf16.err
