# runs the round-2 profiling passes, one gpurun call (= one ncu invocation) each
cd /root/repo
R=tools/gpurun_retry.sh
$R gpurun_out/r2prof_l.log --timeout 900 -- bash tools/r2_prof_launches.sh
$R gpurun_out/r2prof_1.log --timeout 600 -- bash tools/r2_prof_one.sh r2_pconv2_res pconv2_kernel python tools/prof_conv.py res 3
$R gpurun_out/r2prof_2.log --timeout 600 -- bash tools/r2_prof_one.sh r2_igemm_down igemm_kernel python tools/prof_conv.py down 3
$R gpurun_out/r2prof_3.log --timeout 600 -- bash tools/r2_prof_one.sh r2_wgrad_res wgrad_kernel python tools/prof_conv.py res_wgrad 3
$R gpurun_out/r2prof_4.log --timeout 600 -- bash tools/r2_prof_one.sh r2_norm_bwd_reg norm_bwd_reg python tools/prof_norm.py 128 32 1 1 2
$R gpurun_out/r2prof_5.log --timeout 600 -- bash tools/r2_prof_one.sh r2_lean_fwd_apply lean_fwd_apply python tools/prof_norm.py 64 64 0 0 2
$R gpurun_out/r2prof_6.log --timeout 600 -- bash tools/r2_prof_one.sh r2_pconv_c7out pconv_kernel python tools/prof_conv.py c7out 3
