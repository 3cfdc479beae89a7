set -x
mkdir -p gpurun_out
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 --no-baselines "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; cut -c1-200 gpurun_out/$2.json; tail -2 gpurun_out/$2.err; }
run 29641 r2_bench_${N}gpu
run 29642 r2_bench_${N}gpu_nosyncbn --no-sync-bn
DTG_SYMM_BN=0 run 29643 r2_bench_${N}gpu_ncclbn
run 29644 r2_bench_stoch128_${N}gpu --workload stoch128
run 29645 r2_bench_stoch256_${N}gpu --workload stoch256
