# usage: bash scripts/gpu_ncu_gemms.sh TAG [extra env]  -- ncu --set full on the four FPT GEMMs of the second block (qkv, proj, fc1, fc2)
TAG=${1:-gemms}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 4 -f -o gpurun_out/${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32768 > gpurun_out/${TAG}.log 2>&1
tail -2 gpurun_out/${TAG}.log | cut -c1-300
