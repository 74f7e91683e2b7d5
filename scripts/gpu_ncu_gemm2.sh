TAG=${1:-ncu2}
mkdir -p gpurun_out
for sh in proj fc2 fc1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/${TAG}_${sh} python scripts/gemm_bench.py --cg 2 --shapes $sh --iters 1 > gpurun_out/${TAG}_${sh}.log 2>&1
  tail -1 gpurun_out/${TAG}_${sh}.log
done
python scripts/gemm_bench.py --cg 2
