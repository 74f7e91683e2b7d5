# per-config evidence (BASELINE.json configs 2-4 + the "chosen" architecture): same bench keys for every architecture
mkdir -p gpurun_out
TAG=${1:-r2cfg}
: > gpurun_out/${TAG}.jsonl
run() { timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --parity-poses 128 "$@" >> gpurun_out/${TAG}.jsonl 2>> gpurun_out/${TAG}.err; echo "rc=$? $*"; }
run --arch hm0
run --arch chosen
run --arch cmu0 --views 5
run --arch cmu0 --views 2
run --arch chosen --views 5 --depth 2
for v in 2 3 5 6 7 8; do run --arch hm0 --views $v --batch 32768; done
for v in 2 4 8; do run --arch kptok --views $v --batch 32768; done
run --arch hm0 --precision tf32 --batch 32768
python - <<PY
import json
for ln in open("gpurun_out/${TAG}.jsonl"):
    try:
        d=json.loads(ln)
    except Exception: continue
    r=d.get("roofline") or {}
    print(d["config"]["arch"], "V",d["config"]["views"], d["dtype"], "B",d["config"]["batch_per_gpu"], "poses/s", round(d["value"]), "TF", round(d["whole_path_tflops"]), "gemm frac", r.get("frac") and round(r["frac"],3), "err", d["parity"] and round(d["parity"]["max_abs_err_over_scale"],5))
    print("    ", {k: round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
PY
