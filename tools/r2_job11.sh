set -x
mkdir -p gpurun_out
for v in 1 2; do
for args in "128 32 1 1 10 80 cin" "128 32 0 0 10 80 in" "64 32 1 0 10 80 in" "256 16 0 0 10 80 in"; do
DTG_REG_VAR=$v timeout 120 python tools/prof_norm.py $args 2>&1 | grep bwd
done
done
for slab in 128 64 32; do
for args in "64 64 0 0 10 80 cin" "32 64 0 0 10 80 cin" "32 64 0 0 10 80 in"; do
DTG_FUSED_SLAB=$slab timeout 120 python tools/prof_norm.py $args 2>&1 | grep "norm "
done
done
DTG_NORM_IMPL=3 timeout 120 python tools/prof_norm.py 64 64 0 0 10 80 cin 2>&1 | grep "norm "
DTG_NORM_IMPL=3 timeout 120 python tools/prof_norm.py 32 64 0 0 10 80 cin 2>&1 | grep "norm "
DTG_REG_VAR=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j11_bench_regvar2.json 2> gpurun_out/r2j11_bench_regvar2.err; cut -c1-160 gpurun_out/r2j11_bench_regvar2.json
