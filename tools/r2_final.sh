# Round-2 final validation on one B200: the driver's own sequence (GPU tests, smoke, reference arm, our arm)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/test_ratios.jsonl
timeout 1700 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -5 > gpurun_out/r2_final_tests.log; cat gpurun_out/r2_final_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_final_ref.err; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_final_bench.err; cut -c1-200 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_final_bench.err
