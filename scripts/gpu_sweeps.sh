# usage: bash scripts/gpu_sweeps.sh TAG  -- BASELINE.json configs 3/4 on one GPU: CMU configs and the V = 2..8 sweep in both FPT token layouts
TAG=${1:-sweep}
mkdir -p gpurun_out
out=gpurun_out/${TAG}.jsonl; : > $out
for v in 2 5; do python -m openmpl_b200.evaluate --arch cmu0 --views $v --poses 524288 2>/dev/null | tail -1 >> $out; done
python -m openmpl_b200.evaluate --arch chosen --views 4 --poses 524288 2>/dev/null | tail -1 >> $out
for v in 2 3 4 5 6 7 8; do
  python -m openmpl_b200.evaluate --arch hm0 --views $v --poses 262144 2>/dev/null | tail -1 >> $out
  python -m openmpl_b200.evaluate --arch kptok --views $v --poses 262144 2>/dev/null | tail -1 >> $out
done
python - <<PY
import json
for ln in open("$out"):
    d=json.loads(ln); print(d["config"]["workload"], "| poses/s", round(d["value"]), "| TF/s", round(d["tflops"],1), "| mpjpe", round(d["mpjpe_cm"]["absolute"],2))
PY
