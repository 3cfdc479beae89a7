timeout 120 python tools/infer_sweep.py > gpurun_out/infer_sweep.json 2> gpurun_out/infer_sweep.err; cat gpurun_out/infer_sweep.json; tail -2 gpurun_out/infer_sweep.err
for W in stoch128 stoch256; do
  timeout 200 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$W.json')); print('$W', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('torch_gpu_baseline'))"
done
