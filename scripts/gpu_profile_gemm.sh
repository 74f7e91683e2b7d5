TAG=${1:-prof}
mkdir -p gpurun_out
python scripts/gemm_bench.py --cg 1 > gpurun_out/${TAG}_gemm_cg1.txt 2>&1
python scripts/gemm_bench.py --cg 2 > gpurun_out/${TAG}_gemm_cg2.txt 2>&1
cat gpurun_out/${TAG}_gemm_cg1.txt gpurun_out/${TAG}_gemm_cg2.txt
for cg in 1 2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/${TAG}_gemm_cg${cg} python scripts/gemm_bench.py --cg $cg --shapes qkv --iters 1 > gpurun_out/${TAG}_ncu_cg${cg}.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_cg${cg}.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_bench_under_ncu.log | cut -c1-300
ls -la gpurun_out | tail -12
