set -x
mkdir -p gpurun_out
ncu --query-metrics --chip gb100 > gpurun_out/r2_ncu_query_metrics.txt 2>&1 || ncu --query-metrics > gpurun_out/r2_ncu_query_metrics.txt 2>&1
grep -i -c "" gpurun_out/r2_ncu_query_metrics.txt
grep -i "tensor\|tmem\|umma\|utc" gpurun_out/r2_ncu_query_metrics.txt | head -80
