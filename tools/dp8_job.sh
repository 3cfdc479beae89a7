PORT=29811
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "rc=$?"; cut -c1-330 gpurun_out/bench_8gpu.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_8gpu.err | tail -5 | cut -c1-300
