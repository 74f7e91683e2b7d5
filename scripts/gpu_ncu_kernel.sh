# usage: bash scripts/gpu_ncu_kernel.sh TAG KERNEL_REGEX [skip]   -- ncu --set full on one launch of a kernel inside a short bench forward
TAG=${1:-k}; KREGEX=${2:-spt_fused}; SKIP=${3:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP} -c 1 -f -o gpurun_out/${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32768 > gpurun_out/${TAG}.log 2>&1
tail -2 gpurun_out/${TAG}.log | cut -c1-200
