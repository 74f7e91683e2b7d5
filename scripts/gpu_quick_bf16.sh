# usage: bash scripts/gpu_quick_bf16.sh [tag]  -- bf16 parity tests + bf16 bench (fast iteration on a kernel change)
TAG=${1:-q}
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_parity_gpu.py -q -m gpu --timeout 300 -k "bf16 or fused_spt or chunking or packed or layernorm_fusion" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/${TAG}_rc.txt
timeout -k 5 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bench rc=$?" >> gpurun_out/${TAG}_rc.txt
cat gpurun_out/${TAG}_rc.txt; tail -15 gpurun_out/${TAG}_tests.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_bf16.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "roof", d["roofline"]["achieved"], d["roofline"]["frac"], "parity", d["parity"])
        for k,v in d["breakdown"].items(): print("   ", k, round(v["ms_per_step"],3), v["launches_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/${TAG}_bench_bf16.err
