set -x
mkdir -p gpurun_out
timeout 600 python tools/step_profile.py > gpurun_out/r2_step_profile_graph_timed.txt 2> gpurun_out/r2j26.err; tail -3 gpurun_out/r2j26.err; head -60 gpurun_out/r2_step_profile_graph_timed.txt | cut -c1-250
