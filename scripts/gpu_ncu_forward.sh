# usage: bash scripts/gpu_ncu_forward.sh TAG   -- ncu --set full on the first 10 launches of one forward (embed, SPT, token
# build, LN, QKV GEMM, attention, proj GEMM, LN, fc1 GEMM, fc2 GEMM) + the head kernel, batch 32768 (one chunk = 95 launches)
TAG=${1:-fwd}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -s 190 -c 10 -f -o gpurun_out/${TAG}_first10 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32768 > gpurun_out/${TAG}_first10.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:head_fused -s 2 -c 1 -f -o gpurun_out/${TAG}_head python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32768 > gpurun_out/${TAG}_head.log 2>&1
tail -2 gpurun_out/${TAG}_first10.log | cut -c1-300
ls -la gpurun_out/${TAG}*
