cd /root/repo
R=tools/gpurun_retry.sh
PROF_HEAD=2 $R gpurun_out/r2prof_7.log --timeout 600 -- env PROF_HEAD=2 bash tools/r2_prof_one.sh r2_tail7_c7out tail7_kernel python tools/prof_conv.py c7out 3
$R gpurun_out/r2prof_8.log --timeout 600 -- env PROF_FLAT=1 bash tools/r2_prof_one.sh r2_pconv2_res_dgrad_flat pconv2_kernel python tools/prof_conv.py res_dgrad 3
