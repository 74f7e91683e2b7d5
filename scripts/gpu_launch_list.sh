# usage: bash scripts/gpu_launch_list.sh TAG  -- every kernel launch of one bench step with its device time (ncu, cold-cache, serialised)
TAG=${1:-launches}
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}.log 2>&1
wc -l gpurun_out/${TAG}.csv
