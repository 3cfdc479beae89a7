# Round-1 (second half) evidence: plain bench first, then the ncu launch list of the same command, then one
# `ncu --set full` capture per heavy kernel (each micro-driver runs plain before it runs under ncu).
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err
python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/r1b_bench_prof_plain.json 2> gpurun_out/r1b_bench_prof_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches_full.csv python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/r1b_bench_under_ncu.log 2>&1
python tools/timeline.py --out gpurun_out/r1b_timeline.json > gpurun_out/r1b_timeline.txt 2>&1
python tools/prof_conv.py res 3 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:pconv2_kernel -c 1 -s 2 -o gpurun_out/r1b_pconv2_res -f python tools/prof_conv.py res 3 > gpurun_out/ncu_r1b_res.log 2>&1
python tools/prof_conv.py res_dgrad 3 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:pconv2_kernel -c 1 -s 2 -o gpurun_out/r1b_pconv2_res_dgrad -f python tools/prof_conv.py res_dgrad 3 > gpurun_out/ncu_r1b_res_dgrad.log 2>&1
python tools/prof_conv.py res_wgrad 3 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:wgrad_kernel -c 1 -s 2 -o gpurun_out/r1b_wgrad_res -f python tools/prof_conv.py res_wgrad 3 > gpurun_out/ncu_r1b_wgrad.log 2>&1
python tools/prof_conv.py c7out 3 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:pconv_kernel -c 1 -s 2 -o gpurun_out/r1b_pconv_c7out -f python tools/prof_conv.py c7out 3 > gpurun_out/ncu_r1b_c7out.log 2>&1
python tools/prof_norm.py 128 32 1 1 2 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:norm_bwd_reg -c 1 -s 2 -o gpurun_out/r1b_norm_bwd_reg -f python tools/prof_norm.py 128 32 1 1 2 > gpurun_out/ncu_r1b_normreg.log 2>&1
python tools/prof_norm.py 64 64 0 0 2 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:norm_fwd_fused -c 1 -s 2 -o gpurun_out/r1b_norm_fwd_fused -f python tools/prof_norm.py 64 64 0 0 2 > gpurun_out/ncu_r1b_normfwd.log 2>&1
ls -la gpurun_out | grep r1b
