set -x
mkdir -p gpurun_out
DTG_BENCH_NO_KERNEL_PROFILE=1 python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/r2_bench_prof_plain.json 2> gpurun_out/r2_bench_prof_plain.err || exit 1
cut -c1-200 gpurun_out/r2_bench_prof_plain.json
DTG_BENCH_NO_KERNEL_PROFILE=1 timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_full.csv python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_launches_full.csv
