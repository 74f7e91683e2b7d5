# usage: bash scripts/gpu_8x.sh TAG N  -- bench.py and the 10 M-pose evaluation (BASELINE.json config 5) on N GPUs of one box
TAG=${1:-x8}; N=${2:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --parity-poses 64 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 -m openmpl_b200.evaluate --arch hm0 --views 4 --poses 10000000 > gpurun_out/${TAG}_eval10m.json 2> gpurun_out/${TAG}_eval10m.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 -m openmpl_b200.evaluate --arch cmu0 --views 5 --poses 8388608 > gpurun_out/${TAG}_evalcmu.json 2> gpurun_out/${TAG}_evalcmu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus $N --steps 10 --warmup 3 --parity-poses 64 --arch cmu0 --views 5 > gpurun_out/${TAG}_bench_cmu5.json 2> gpurun_out/${TAG}_bench_cmu5.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus $N --steps 10 --warmup 3 --parity-poses 64 --arch chosen > gpurun_out/${TAG}_bench_chosen.json 2> gpurun_out/${TAG}_bench_chosen.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29526 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "reference arm under torchrun rc=$?"; tail -c 600 gpurun_out/${TAG}_ref.json
python - <<PY
import json
for f in ("bench","bench_cmu5","bench_chosen","eval10m","evalcmu"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","n_gpus","ms_per_step","ms_total","poses","mpjpe_cm")}, d.get("e2e",{}).get("value"))
    except Exception as e:
        print(f,"ERR",e); print(open("gpurun_out/${TAG}_%s.err"%f).read()[-800:])
PY
