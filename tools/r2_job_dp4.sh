set -x
mkdir -p gpurun_out
N=4
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 --no-baselines "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; cut -c1-200 gpurun_out/$2.json; tail -2 gpurun_out/$2.err; }
run 29651 r2_bench_4gpu
run 29654 r2_bench_stoch128_4gpu --workload stoch128
run 29655 r2_bench_stoch256_4gpu --workload stoch256
