set -x
mkdir -p gpurun_out
{
PROF_KERNELS=1 timeout 120 python tools/prof_conv.py res_wgrad 10
DTG_WGRAD_DBG=3 PROF_KERNELS=1 timeout 120 python tools/prof_conv.py res_wgrad 10
PROF_KERNELS=1 timeout 120 python tools/prof_conv.py c3a_wgrad 10
PROF_KERNELS=1 timeout 120 python tools/prof_conv.py c7in_wgrad 10
PROF_KERNELS=1 timeout 120 python tools/prof_norm.py 128 32 1 1
DTG_DEBUG_OCC=1 PROF_KERNELS=1 timeout 120 python tools/prof_norm.py 128 16 0 0 10 160
} > gpurun_out/r2j2_kernels.log 2>&1; cat gpurun_out/r2j2_kernels.log
python tools/prof_norm.py 128 32 1 1 2 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:norm_bwd_tma -c 1 -s 2 -o gpurun_out/r2j2_norm_bwd_tma -f python tools/prof_norm.py 128 32 1 1 2 > gpurun_out/r2j2_ncu_normbwd.log 2>&1
python tools/prof_conv.py res_wgrad 3 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:wgrad_reduce -c 1 -s 2 -o gpurun_out/r2j2_wgrad_reduce -f python tools/prof_conv.py res_wgrad 3 > gpurun_out/r2j2_ncu_reduce.log 2>&1
ls -la gpurun_out | grep r2j2
