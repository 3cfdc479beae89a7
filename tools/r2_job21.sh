set -x
timeout 300 python tools/perm_noise_oracle.py --kind uniform 2>&1 | tail -8 | grep -v bf16
timeout 300 python tools/perm_noise_oracle.py --kind uniform --batch 32 2>&1 | tail -8 | grep -v bf16
timeout 300 python tools/perm_noise_oracle.py --wscale 5 2>&1 | tail -8 | grep -v bf16
timeout 300 python tools/perm_noise_oracle.py --wscale 5 --kind uniform --batch 16 2>&1 | tail -8 | grep -v bf16
