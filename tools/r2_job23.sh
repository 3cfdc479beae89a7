set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -4
PROF_KERNELS=1 timeout 120 python tools/prof_conv.py c7in_dgrad 10 2>&1 | tail -4
