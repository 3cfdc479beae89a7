set -x
mkdir -p gpurun_out
for v in 0 1; do
for args in "128 32 1 1 10 80 cin" "128 32 0 0 10 80 in" "64 32 1 0 10 80 in" "256 16 0 0 10 80 in"; do
DTG_REG_VAR=$v timeout 120 python tools/prof_norm.py $args 2>&1 | grep bwd
done
done
DTG_REG_VAR=1 timeout 300 python -m pytest tests/test_norm_gpu.py -q -x 2>&1 | tail -2
DTG_REG_VAR=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j10_bench_regvar.json 2> gpurun_out/r2j10_bench_regvar.err; cut -c1-160 gpurun_out/r2j10_bench_regvar.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines --workload stoch128 > gpurun_out/r2_bench_stoch128_1gpu.json 2> gpurun_out/r2j10_s128.err; cut -c1-160 gpurun_out/r2_bench_stoch128_1gpu.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines --workload stoch256 > gpurun_out/r2_bench_stoch256_1gpu.json 2> gpurun_out/r2j10_s256.err; cut -c1-160 gpurun_out/r2_bench_stoch256_1gpu.json
