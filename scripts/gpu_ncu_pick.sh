# usage: bash scripts/gpu_ncu_pick.sh TAG REGEX SKIP COUNT  -- ncu --set full on COUNT launches matching REGEX after SKIP matches (batch 32768)
TAG=$1; RE=$2; SKIP=${3:-1}; CNT=${4:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${RE} -s ${SKIP} -c ${CNT} -f -o gpurun_out/${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32768 > gpurun_out/${TAG}.log 2>&1
tail -1 gpurun_out/${TAG}.log | cut -c1-200
