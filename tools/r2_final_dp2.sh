set -x
mkdir -p gpurun_out
rm -f gpurun_out/test_ratios.jsonl
timeout 900 python -m pytest tests/test_dp_gpu.py -q --tb=short 2>&1 | tail -8
