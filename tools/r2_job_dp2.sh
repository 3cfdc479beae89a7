set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dp_gpu.py -q --tb=short -rf 2>&1 | tail -15 > gpurun_out/r2dp2_tests.log; cat gpurun_out/r2dp2_tests.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --no-baselines "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; cut -c1-220 gpurun_out/$2.json; }
run 29611 r2_bench_2gpu
run 29612 r2_bench_2gpu_nosyncbn --no-sync-bn
run 29613 r2_bench_stoch128_2gpu --workload stoch128
run 29614 r2_bench_stoch256_2gpu --workload stoch256
