set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_networks_gpu.py tests/test_step_gpu.py -q --tb=short -k "extended or n_blocks or wrong_grid or enc_A_B or latent_nets or cin_generator" 2>&1 | tail -30
