set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x --tb=short 2>&1 | tail -12
