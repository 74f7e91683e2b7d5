mkdir -p gpurun_out
export PROBE_CASES=hm0_v4_d12,chosen_v4_d12,cmu0_v2_d2,cmu_v5_d2_hm0flags,cmu_v5_d2_chosen,kptok_v4_d12,sweep_viewtok_v8
timeout -k 5 600 python scripts/tf32_error_probe.py tf32 bf16 > gpurun_out/r2c_probe.log 2>&1; tail -20 gpurun_out/r2c_probe.log
timeout -k 5 600 python bench.py --precision tf32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_tf32.json 2> gpurun_out/r2c_bench_tf32.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c_bench_tf32.json").read().strip().splitlines()[-1])
    print("tf32 value", round(d["value"]), "ms", round(d["ms_per_step"],2), "parity", d["parity"])
    print("   ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2c_bench_tf32.err").read()[-1500:])
PY
timeout -k 5 1500 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2c_tests.log
