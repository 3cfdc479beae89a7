set -x
mkdir -p gpurun_out
for cap in 112 100; do
DTG_SMEM_CAP_KB=$cap timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j19_bench_cap$cap.json 2> gpurun_out/r2j19_bench_cap$cap.err; cut -c1-160 gpurun_out/r2j19_bench_cap$cap.json; tail -2 gpurun_out/r2j19_bench_cap$cap.err
done
