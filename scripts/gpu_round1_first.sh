mkdir -p gpurun_out; nvidia-smi > gpurun_out/smi.txt 2>&1
timeout -k 5 900 python -m pytest tests/test_parity_gpu.py -q -k "fp32_path or native or cpu_inputs or repacked" --timeout 300 > gpurun_out/t_fp32.log 2>&1; echo "fp32 rc=$?" >> gpurun_out/rc.txt
timeout -k 5 300 python -m pytest tests/test_metric_inputs_gpu.py -q --timeout 120 > gpurun_out/t_metric.log 2>&1; echo "metric rc=$?" >> gpurun_out/rc.txt
timeout -k 5 600 python -m pytest tests/test_gemm_gpu.py -q -k "cg1 or rejects" --timeout 120 > gpurun_out/t_gemm1.log 2>&1; echo "gemm1 rc=$?" >> gpurun_out/rc.txt
timeout -k 5 600 python -m pytest tests/test_gemm_gpu.py -q -k cg2 --timeout 120 > gpurun_out/t_gemm2.log 2>&1; echo "gemm2 rc=$?" >> gpurun_out/rc.txt
timeout -k 5 1200 python -m pytest tests/test_parity_gpu.py -q -k "not fp32_path" --timeout 300 > gpurun_out/t_rest.log 2>&1; echo "rest rc=$?" >> gpurun_out/rc.txt
timeout -k 5 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; for f in t_fp32 t_metric t_gemm1 t_gemm2 t_rest; do echo "== $f"; tail -5 gpurun_out/$f.log; done; cat gpurun_out/bench_bf16.json; tail -3 gpurun_out/bench_bf16.err
