set -x
for bt in 16 32 64; do
for extra in "" "--no-sync-bn"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tools/dp_parity.py --precision tf32 --batch $bt $extra 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['batch'], d['sync_bn'], 'grad', {k:round(v,5) for k,v in d['grad_rel'].items()}, 'perm', {k:round(v,5) for k,v in d['perm_noise_grad_rel'].items()})"
done
done
