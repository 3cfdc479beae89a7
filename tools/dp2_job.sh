# 2-rank data-parallel check: the driver's launch line under a timeout, then the same with lanes disabled for comparison
PORT=29711
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_lanes.json 2> gpurun_out/bench_2gpu_lanes.err
echo "rc=$?"; cut -c1-420 gpurun_out/bench_2gpu_lanes.json; tail -5 gpurun_out/bench_2gpu_lanes.err | cut -c1-300
