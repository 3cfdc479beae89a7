set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stoch_gpu.py tests/test_norm_gpu.py -q -x --tb=short -k "stoch_train_instance or loss_fused" 2>&1 | tail -6
for i in 1 2 3; do timeout 300 python bench.py --steps 50 --warmup 5 --no-baselines > gpurun_out/r2_bench_repeat$i.json 2> gpurun_out/r2j20_$i.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_repeat$i.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"; done
