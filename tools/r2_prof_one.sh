# usage: bash tools/r2_prof_one.sh <tag> <kernel regex> <extra ncu args or ''> -- <plain command ...>
set -x
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
"$@" > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -2 gpurun_out/plain_$TAG.log
TM=sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum.per_second,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__sass_inst_executed_op_utcmma.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum,sm__inst_executed_pipe_tc.sum
ncu --set full --metrics $TM --import-source on --clock-control none -k regex:$KRE -c 1 -s 2 -o gpurun_out/$TAG -f "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/$TAG.ncu-rep
