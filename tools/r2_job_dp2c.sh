set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -q --tb=short -rf 2>&1 | tail -8 > gpurun_out/r2dp2c_tests.log; cat gpurun_out/r2dp2c_tests.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --no-baselines "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; cut -c1-200 gpurun_out/$2.json; tail -2 gpurun_out/$2.err; }
run 29631 r2_bench_2gpu_symm
DTG_SYMM_BN=0 run 29632 r2_bench_2gpu_ncclbn
