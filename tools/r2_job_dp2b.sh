run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dp_parity.py "${@:2}" 2>/dev/null | tail -1; }
run 29621 --precision bf16 --batch 8
run 29622 --precision tf32 --batch 64
run 29623 --precision tf32 --batch 8 --graph
