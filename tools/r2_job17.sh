set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x --tb=short -k "tail_filter or nchw_tanh" 2>&1 | tail -6
for d in 0 4; do DTG_T7_DBG=$d PROF_HEAD=2 timeout 120 python tools/prof_conv.py c7out 10 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_networks_gpu.py tests/test_step_gpu.py tests/test_stoch_gpu.py -q -x --tb=short 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j17_bench.json 2> gpurun_out/r2j17_bench.err; cut -c1-160 gpurun_out/r2j17_bench.json
