# round 2, GPU call 4: co-residency experiment (shared-memory budget of the tensor-core CTAs), full test ratios, new tests
set -x
mkdir -p gpurun_out
for cap in 227 204 180 160; do
  DTG_SMEM_CAP_KB=$cap timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j4_bench_cap$cap.json 2> gpurun_out/r2j4_bench_cap$cap.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2j4_bench_cap$cap.json"))
    print("cap $cap:", d["ms_per_step"], d["value"], d["e2e"]["value"], (d.get("roofline") or {}).get("serial_step_us"))
except Exception as e:
    print("cap $cap failed", e)
PY
done
tail -3 gpurun_out/r2j4_bench_cap227.err
rm -f gpurun_out/test_ratios.jsonl
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2j4_tests_all.log; cat gpurun_out/r2j4_tests_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
