# usage: bash scripts/gpu_sanitize.sh TAG  -- compute-sanitizer memcheck + racecheck of small forwards (every precision, graph + eager)
TAG=${1:-san}
mkdir -p gpurun_out
timeout -k 5 420 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/${TAG}_memcheck.log
timeout -k 5 420 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_race.py > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -c "Error: Race" gpurun_out/${TAG}_racecheck.log; grep "Error: Race" gpurun_out/${TAG}_racecheck.log | cut -c1-150 | sort | uniq -c | head -20; tail -3 gpurun_out/${TAG}_racecheck.log
