set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x --tb=short -k "tail_filter or nchw_tanh" 2>&1 | tail -12
for d in 0 1 2 4 3 7; do DTG_P2_DBG=$d timeout 120 python tools/prof_conv.py res 10 2>&1 | tail -1; done
PROF_HEAD=1 timeout 120 python tools/prof_conv.py c7out 10 2>&1 | tail -1
PROF_HEAD=2 timeout 120 python tools/prof_conv.py c7out 10 2>&1 | tail -1
