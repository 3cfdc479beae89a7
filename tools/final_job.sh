# round-end check, the driver's order: GPU tests, smoke, bench (both arms)
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/final_tests.log; cat gpurun_out/final_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --no-baselines > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; wc -l gpurun_out/final_bench.json; cut -c1-200 gpurun_out/final_bench.json
python - <<PY
import json
d=json.load(open("gpurun_out/final_bench.json"))
print({k: d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks")})
print(d["roofline"]["frac"], d["roofline"]["tensor_kernels_share_of_step"])
PY
