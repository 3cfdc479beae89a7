set -x
mkdir -p gpurun_out
rm -f gpurun_out/test_ratios.jsonl
timeout 1500 python -m pytest tests -q -m gpu --tb=line -rf -x 2>&1 | tail -6 > gpurun_out/r2j9_tests_all.log; cat gpurun_out/r2j9_tests_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j9_bench.json 2> gpurun_out/r2j9_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2j9_bench.json
for cap in 200 180; do
DTG_SMEM_CAP_KB=$cap timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j9_bench_cap$cap.json 2> gpurun_out/r2j9_bench_cap$cap.err; cut -c1-160 gpurun_out/r2j9_bench_cap$cap.json
done
