set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_norm_gpu.py tests/test_wgrad_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r2j3_tests_a.log; cat gpurun_out/r2j3_tests_a.log
{
for args in "128 32 1 1" "128 32 0 0" "64 64 0 0" "32 64 0 0" "128 16 0 0 10 160" "256 15 0 0 10 160"; do
  DTG_DEBUG_OCC=1 timeout 120 python tools/prof_norm.py $args
done
DTG_NO_TMA_NORM=1 timeout 120 python tools/prof_norm.py 128 32 1 1
for c in res_wgrad c3a_wgrad; do timeout 120 python tools/prof_conv.py $c 10; done
} > gpurun_out/r2j3_micro.log 2>&1; grep -v "^+" gpurun_out/r2j3_micro.log
rm -f gpurun_out/test_ratios.jsonl
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2j3_tests_all.log; cat gpurun_out/r2j3_tests_all.log
cat gpurun_out/test_ratios.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j3_bench.json 2> gpurun_out/r2j3_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2j3_bench.json
DTG_NO_TMA_NORM=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j3_bench_oldnorm.json 2> gpurun_out/r2j3_bench_oldnorm.err; cut -c1-300 gpurun_out/r2j3_bench_oldnorm.json
