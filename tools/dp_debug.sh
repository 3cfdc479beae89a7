# diagnostic: 2-rank bench under short timeouts, four configurations
run() { echo "=== $*"; env "$@" DTG_BENCH_VERBOSE=50 timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 3 --warmup 3 --no-baselines $EXTRA 2>&1 | grep -E "bench rank|metric|Error|error|File \"/root/repo|line [0-9]+ in" | cut -c1-220 | tail -25; PORT=$((PORT+1)); }
PORT=29611
EXTRA="--no-graph" run DTG_NO_PDL=1
EXTRA="--no-graph" run DTG_X=1
EXTRA="" run DTG_NO_PDL=1
EXTRA="" run DTG_X=1
