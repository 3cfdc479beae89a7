set -x
mkdir -p gpurun_out
rm -f gpurun_out/test_ratios.jsonl
timeout 1700 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -8 > gpurun_out/r2j18_tests_all.log; cat gpurun_out/r2j18_tests_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
