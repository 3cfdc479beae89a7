# usage: bash tools/perf_job.sh TAG [pytest files...]   -- parity tests, per-op profile and a short bench for one change
TAG=$1; shift
timeout 400 python -m pytest "$@" -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 300 python tools/step_profile.py --json gpurun_out/step_profile_${TAG}.json > gpurun_out/step_profile_${TAG}.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/${TAG}_tests.log; head -18 gpurun_out/step_profile_${TAG}.txt | cut -c1-150; cut -c1-330 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
