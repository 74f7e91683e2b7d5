# usage: bash scripts/gpu_ncu_final.sh TAG -- launch list of one bench step + ncu --set full of the view attention and the head
TAG=${1:-final}
mkdir -p gpurun_out
bash scripts/gpu_launch_list.sh ${TAG}_launches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_views -s 2 -c 1 -f -o gpurun_out/${TAG}_attn python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32768 > gpurun_out/${TAG}_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_block -s 1 -c 1 -f -o gpurun_out/${TAG}_head python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32768 > gpurun_out/${TAG}_head.log 2>&1
ls -la gpurun_out/${TAG}_*
