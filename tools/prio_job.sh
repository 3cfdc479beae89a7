for P in 0,0 0,-1 -1,0 0,-2; do
  DTG_LANE_PRIO=$P timeout 200 python bench.py --steps 20 --warmup 5 --no-baselines 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$P', d['ms_per_step'], d['e2e']['value'])"
done
timeout 200 python -m pytest tests/test_stoch_gpu.py -x -q -m gpu -k checkpoint 2>&1 | tail -3
