TAG=${1:-q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -x > gpurun_out/${TAG}_gemmtests.log 2>&1; tail -2 gpurun_out/${TAG}_gemmtests.log
python scripts/gemm_bench.py --cg 1 2>&1 | tee gpurun_out/${TAG}_gemm_cg1.txt
python scripts/gemm_bench.py --cg 2 2>&1 | tee gpurun_out/${TAG}_gemm_cg2.txt
