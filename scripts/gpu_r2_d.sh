mkdir -p gpurun_out
timeout -k 5 180 python -m pytest tests/test_gemm_ln_gpu.py -q -x --timeout 120 > gpurun_out/r2d_gemm_ln.log 2>&1; echo "gemm_ln rc=$?"; tail -25 gpurun_out/r2d_gemm_ln.log
timeout -k 5 300 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 120 > gpurun_out/r2d_gemm.log 2>&1; echo "gemm rc=$?"; tail -5 gpurun_out/r2d_gemm.log
timeout -k 5 600 python -m pytest tests/test_parity_gpu.py -q -x --timeout 120 -k "bf16 or layernorm_fusion" > gpurun_out/r2d_parity.log 2>&1; echo "parity rc=$?"; tail -8 gpurun_out/r2d_parity.log
timeout -k 5 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
    print("bf16 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "roof", d["roofline"]["frac"], "parity", d["parity"])
    print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2d_bench.err").read()[-1500:])
PY
