set -x
python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/r1_bench_prof_plain.json 2> gpurun_out/r1_bench_prof_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_full.csv python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/r1_bench_under_ncu.log 2>&1
python tools/prof_conv.py res 3 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:igemm_kernel -c 1 -s 2 -o gpurun_out/r1_igemm_res -f python tools/prof_conv.py res 3 > gpurun_out/ncu_igemm.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:wgrad_kernel -c 1 -s 2 -o gpurun_out/r1_wgrad_res -f python tools/prof_conv.py res_wgrad 3 > gpurun_out/ncu_wgrad.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:pconv_kernel -c 1 -s 2 -o gpurun_out/r1_pconv_c3a -f python tools/prof_conv.py c3a 3 > gpurun_out/ncu_pconv.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:norm_bwd_reg -c 1 -s 2 -o gpurun_out/r1_norm_bwd_reg -f python tools/prof_norm.py 128 32 1 1 2 > gpurun_out/ncu_normreg.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:norm_fwd_fused -c 1 -s 2 -o gpurun_out/r1_norm_fwd_fused -f python tools/prof_norm.py 64 64 0 0 2 > gpurun_out/ncu_normfwd.log 2>&1
ls -la gpurun_out | tail -12
