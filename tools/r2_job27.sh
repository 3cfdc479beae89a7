set -x
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/r2j27_bench$i.json 2> gpurun_out/r2j27.err; python -c "
import json; d=json.load(open('gpurun_out/r2j27_bench$i.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; done
timeout 300 python bench.py --steps 100 --warmup 5 --no-baselines > gpurun_out/r2j27_bench3.json 2> gpurun_out/r2j27.err; python -c "
import json; d=json.load(open('gpurun_out/r2j27_bench3.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
