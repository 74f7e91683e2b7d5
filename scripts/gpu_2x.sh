# usage: bash scripts/gpu_2x.sh TAG [N]  -- bench.py and a 2 M-pose evaluation on N (default 2) GPUs of one box
TAG=${1:-x2}; N=${2:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 -m openmpl_b200.evaluate --arch hm0 --views 4 --poses 2097152 > gpurun_out/${TAG}_eval.json 2> gpurun_out/${TAG}_eval.err
python - <<PY
import json
for f in ("bench","eval"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","n_gpus","ms_per_step","ms_total","poses","mpjpe_cm")}, d.get("e2e",{}).get("value"))
    except Exception as e:
        print(f,"ERR",e); print(open("gpurun_out/${TAG}_%s.err"%f).read()[-800:])
PY
