# usage: bash scripts/gpu_tests_and_bench.sh [tag]   -- full -m gpu suite + bf16 bench (+ optional extras via env)
TAG=${1:-run}
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/${TAG}_rc.txt
timeout -k 5 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bench rc=$?" >> gpurun_out/${TAG}_rc.txt
MPL_GEMM_CTA_GROUP=1 timeout -k 5 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_bf16_cg1.json 2> gpurun_out/${TAG}_bench_bf16_cg1.err; echo "bench cg1 rc=$?" >> gpurun_out/${TAG}_rc.txt
cat gpurun_out/${TAG}_rc.txt; tail -15 gpurun_out/${TAG}_tests.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_bf16.json","gpurun_out/${TAG}_bench_bf16_cg1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "roof", d["roofline"]["achieved"], d["roofline"]["frac"], "parity", d["parity"])
        for k,v in d["breakdown"].items(): print("   ", k, round(v["ms_per_step"],3), v["launches_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/${TAG}_bench_bf16.err
