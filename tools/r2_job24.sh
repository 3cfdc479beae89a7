set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -4
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j24_bench$i.json 2> gpurun_out/r2j24_bench.err; cut -c1-160 gpurun_out/r2j24_bench$i.json; done
