timeout 400 python -m pytest tests/test_stoch_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/stoch_tests.log; tail -3 gpurun_out/stoch_tests.log
for W in stoch128 stoch256; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$W.json"))
    print("$W", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("torch_gpu_baseline"), d["roofline"]["all_tensor_kernels_tflops"], d["conv_roofline_frac_of_step"])
except Exception as e:
    print("$W failed", e); print(open("gpurun_out/bench_$W.err").read()[-1500:])
PY
done
