mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2e_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 4 -f -o gpurun_out/r2e_gemms python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32768 > gpurun_out/r2e_gemms.log 2>&1
ls -la gpurun_out/r2e_gemms.ncu-rep
