set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_norm_gpu.py tests/test_wgrad_gpu.py -q --tb=line -rf 2>&1 | tail -15 > gpurun_out/r2j5_tests_a.log; cat gpurun_out/r2j5_tests_a.log
{
for c in res_wgrad c3a_wgrad; do PROF_KERNELS=1 DTG_NO_PDL=1 timeout 120 python tools/prof_conv.py $c 10; done
} > gpurun_out/r2j5_micro.log 2>&1; grep -v "^+\|Warn\|_warn" gpurun_out/r2j5_micro.log
rm -f gpurun_out/test_ratios.jsonl
timeout 1500 python -m pytest tests -q -m gpu --tb=line -rf 2>&1 | tail -15 > gpurun_out/r2j5_tests_all.log; cat gpurun_out/r2j5_tests_all.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j5_bench.json 2> gpurun_out/r2j5_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2j5_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j5_bench_ref.json 2> gpurun_out/r2j5_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2j5_bench_ref.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
