set -x
mkdir -p gpurun_out
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 --no-baselines "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; cut -c1-200 gpurun_out/$2.json; tail -2 gpurun_out/$2.err; }
run 29741 r2_bench_final_${N}gpu
