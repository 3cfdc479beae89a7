set -x
mkdir -p gpurun_out
for w in stoch128 stoch256 aug128; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines --workload $w > gpurun_out/r2_bench_final_${w}_1gpu.json 2> gpurun_out/r2j25_$w.err; cut -c1-170 gpurun_out/r2_bench_final_${w}_1gpu.json; tail -1 gpurun_out/r2j25_$w.err
done
