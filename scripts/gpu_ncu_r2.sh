# usage: bash scripts/gpu_ncu_r2.sh TAG  -- ncu --set full of every kernel family of one bf16 forward chunk (32 768 poses)
TAG=${1:-r2ncu}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --parity-poses 0 --batch 32768"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|qkv_attn_kernel" -s 4 -c 4 -f -o gpurun_out/${TAG}_gemms $B > gpurun_out/${TAG}_gemms.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spt_fused -s 1 -c 1 -f -o gpurun_out/${TAG}_spt $B > gpurun_out/${TAG}_spt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"head_block|ln_prep" -s 0 -c 2 -f -o gpurun_out/${TAG}_mem $B > gpurun_out/${TAG}_mem.log 2>&1
# the "chosen" architecture (D = 544, 68-wide heads) and the pose-aligned V = 5 tiling of the fused QKV + attention kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|qkv_attn_kernel" -s 4 -c 4 -f -o gpurun_out/${TAG}_chosen $B --arch chosen > gpurun_out/${TAG}_chosen.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qkv_attn_kernel" -s 1 -c 1 -f -o gpurun_out/${TAG}_v5 $B --arch cmu0 --views 5 > gpurun_out/${TAG}_v5.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep
