mkdir -p gpurun_out
TAG=${1:-r2fa}
timeout -k 5 150 python -m pytest tests/test_qkv_attn_gpu.py -q -x --timeout 60 > gpurun_out/${TAG}_unit.log 2>&1; echo "unit rc=$?"; tail -12 gpurun_out/${TAG}_unit.log
