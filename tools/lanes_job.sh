set -x
timeout 300 python -m pytest tests/test_step_gpu.py tests/test_stoch_gpu.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/lanes_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/bench_lanes.json 2> gpurun_out/bench_lanes.err
tail -3 gpurun_out/lanes_tests.log; cat gpurun_out/bench_lanes.json | cut -c1-400; tail -5 gpurun_out/bench_lanes.err
