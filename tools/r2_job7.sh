set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_norm_gpu.py tests/test_wgrad_gpu.py -q --tb=line -rf 2>&1 | tail -6 > gpurun_out/r2j7_tests_a.log; cat gpurun_out/r2j7_tests_a.log
{
for args in "128 32 1 1" "128 32 0 0" "64 64 0 0" "32 64 0 0" "128 16 0 0 10 160" "256 15 0 0 10 160"; do
  DTG_NORM_IMPL=2 timeout 120 python tools/prof_norm.py $args
  DTG_NORM_IMPL=0 timeout 120 python tools/prof_norm.py $args
done
DTG_NORM_IMPL=2 PROF_KERNELS=1 DTG_NO_PDL=1 timeout 120 python tools/prof_norm.py 128 32 1 1
PROF_KERNELS=1 DTG_NO_PDL=1 timeout 120 python tools/prof_conv.py res_wgrad 10
PROF_KERNELS=1 DTG_NO_PDL=1 timeout 120 python tools/prof_conv.py c3a_wgrad 10
} > gpurun_out/r2j7_micro.log 2>&1; grep -v "^+\|Warn\|_warn" gpurun_out/r2j7_micro.log
timeout 1500 python -m pytest tests -q -m gpu --tb=line -rf -x 2>&1 | tail -8 > gpurun_out/r2j7_tests_all.log; cat gpurun_out/r2j7_tests_all.log
for impl in 2 0; do
DTG_NORM_IMPL=$impl timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j7_bench_impl$impl.json 2> gpurun_out/r2j7_bench_impl$impl.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2j7_bench_impl$impl.json
done
