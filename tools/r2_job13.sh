set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_networks_gpu.py tests/test_norm_gpu.py -q -x --tb=short 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j13_bench.json 2> gpurun_out/r2j13_bench.err; cut -c1-160 gpurun_out/r2j13_bench.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines --workload aug128 > gpurun_out/r2_bench_aug128_1gpu.json 2> gpurun_out/r2j13_aug128.err; cut -c1-160 gpurun_out/r2_bench_aug128_1gpu.json; tail -3 gpurun_out/r2j13_aug128.err
