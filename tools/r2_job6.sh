set -x
mkdir -p gpurun_out
{
for sub in 1 4 8 16; do DTG_REDUCE_SUB=$sub PROF_KERNELS=1 DTG_NO_PDL=1 timeout 120 python tools/prof_conv.py res_wgrad 10; done
for sub in 4 16; do DTG_WGRAD_DBG=8 DTG_REDUCE_SUB=$sub PROF_KERNELS=1 DTG_NO_PDL=1 timeout 120 python tools/prof_conv.py res_wgrad 10; done
DTG_WGRAD_DBG=4 PROF_KERNELS=1 DTG_NO_PDL=1 timeout 120 python tools/prof_conv.py res_wgrad 10
DTG_WGRAD_DBG=7 PROF_KERNELS=1 DTG_NO_PDL=1 timeout 120 python tools/prof_conv.py res_wgrad 10
} > gpurun_out/r2j6_micro.log 2>&1; grep -v "^+\|Warn\|_warn" gpurun_out/r2j6_micro.log
timeout 600 python -m pytest tests/test_step_gpu.py tests/test_norm_gpu.py::test_losses_and_gather -q --tb=line -rf 2>&1 | tail -5
timeout 300 python tools/timeline.py --out gpurun_out/r2j6_timeline.json > gpurun_out/r2j6_timeline.txt 2>&1; head -12 gpurun_out/r2j6_timeline.txt | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j6_bench.json 2> gpurun_out/r2j6_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2j6_bench.json
