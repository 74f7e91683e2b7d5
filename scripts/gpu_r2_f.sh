mkdir -p gpurun_out
TAG=${1:-r2f}
timeout -k 5 120 python -m pytest tests/test_gemm_ln_gpu.py -q -x --timeout 60 > gpurun_out/${TAG}_gemm_ln.log 2>&1; echo "gemm_ln rc=$?"; tail -4 gpurun_out/${TAG}_gemm_ln.log
timeout -k 5 200 python -m pytest tests/test_parity_gpu.py -q -x --timeout 60 -k "bf16 or layernorm_fusion" > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/${TAG}_parity.log
for i in 1 2; do
timeout -k 5 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench$i.json 2> gpurun_out/${TAG}_bench$i.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench$i.json").read().strip().splitlines()[-1])
    print("bf16 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "roof", round(d["roofline"]["frac"],3), "parity", d["parity"])
    print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${TAG}_bench$i.err").read()[-1500:])
PY
done
