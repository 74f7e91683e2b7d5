# round 2, first GPU pass of the split-operand (fp32-grade) mode: GEMM unit tests, error probe on the parity cases, bench
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 300 > gpurun_out/r2a_gemm_tests.log 2>&1; echo "gemm tests rc=$?"
tail -5 gpurun_out/r2a_gemm_tests.log
PROBE_CASES=hm0_v4_d12,chosen_v4_d12,cmu0_v2_d2,cmu_v5_d2_hm0flags,cmu_v5_d2_chosen,sweep_viewtok_v2,sweep_viewtok_v8,kptok_v4_d12,sweep_kptok_v6 timeout -k 5 600 python scripts/tf32_error_probe.py tf32 bf16 > gpurun_out/r2a_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2a_probe.log | tail -30
timeout -k 5 600 python scripts/gemm_bench.py --dtype tf32 --cg 2 > gpurun_out/r2a_gemm_bench.log 2>&1; tail -5 gpurun_out/r2a_gemm_bench.log
timeout -k 5 600 python bench.py --precision tf32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_tf32.json 2> gpurun_out/r2a_bench_tf32.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2a_bench_tf32.json").read().strip().splitlines()[-1])
    print("tf32 value", round(d["value"]), "ms", round(d["ms_per_step"],2), "roof", d["roofline"]["frac"], "parity", d["parity"])
    print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2a_bench_tf32.err").read()[-1500:])
PY
