# usage: gpurun --gpus 2 -- bash scripts/gpu_2x.sh TAG  -- the reference's DataParallel path on two GPUs + a 2-rank bench
TAG=${1:-r2_2gpu}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus.txt
timeout -k 5 300 python -m pytest tests/test_parity_gpu.py -v -k "data_parallel or sweep_viewtok_v3 or sweep_viewtok_v6 or cmu_v5" --timeout 120 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --parity-poses 64 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
    print("2 GPUs: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "mpjpe", d["mpjpe_cm"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
