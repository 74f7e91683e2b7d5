# usage: bash scripts/gpu_ab_lib.sh TAG OLD_LIB "bench args" ["bench args" ...]  -- same-box A/B of two builds of the library (old, new, old, new)
TAG=$1; OLD=$2; shift 2
mkdir -p gpurun_out
NEW=$PWD/openmpl_b200/libmpl_b200.so
for a in "$@"; do
  for rep in 1 2; do
    for l in $OLD $NEW; do
      MPL_B200_LIB=$l timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --parity-poses 64 $a 2> gpurun_out/${TAG}.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$a', '$l'[-26:], 'poses/s', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'err', round(d['parity']['max_abs_err_over_scale'], 5), {k: round(v['ms_per_step'], 2) for k, v in d['breakdown'].items()})"
    done
  done
done
