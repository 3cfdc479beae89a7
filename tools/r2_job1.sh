# round 2, GPU call 1: parity of the new norm / wgrad-reduce kernels, then micro-timings (new vs old), then the full suite + bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_norm_gpu.py tests/test_wgrad_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2j1_tests_a.log; cat gpurun_out/r2j1_tests_a.log
{
for args in "128 32 1 1" "128 32 1 0" "128 32 0 0" "64 64 0 0" "32 64 0 0" "64 64 0 0 10 80 cin" "128 16 0 0 10 160" "256 15 0 0 10 160" "256 14 0 0 10 160"; do
  timeout 120 python tools/prof_norm.py $args
  DTG_NO_TMA_NORM=1 timeout 120 python tools/prof_norm.py $args
done
} > gpurun_out/r2j1_norm.log 2>&1; cat gpurun_out/r2j1_norm.log
{
for c in res_wgrad c3a_wgrad c3b_wgrad c7in_wgrad c7out_wgrad; do timeout 120 python tools/prof_conv.py $c 10; done
for d in 1 2 4 3; do DTG_WGRAD_DBG=$d timeout 120 python tools/prof_conv.py res_wgrad 10; done
for d in 1 2; do DTG_WGRAD_DBG=$d timeout 120 python tools/prof_conv.py c3a_wgrad 10; done
} > gpurun_out/r2j1_wgrad.log 2>&1; cat gpurun_out/r2j1_wgrad.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2j1_tests_all.log; cat gpurun_out/r2j1_tests_all.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j1_bench.json 2> gpurun_out/r2j1_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2j1_bench.json
