"""Selected metrics of `ncu --set full` reports as one JSON (profiles/*_ncu_summary.json).
usage: python tools/ncu_summary.py out.json name=file.ncu-rep[:flops=F|:bytes=B] ..."""
import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "launch__grid_size": "grid",
    "launch__registers_per_thread": "regs",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed_max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__inst_executed_pipe_tensor.sum": "tensor_pipe_insts",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
    # tcgen05-aware counters (ncu --query-metrics on the B200 box, profiles/r2_ncu_tensor_metrics.txt)
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum": "utchmma_bf16_math_ops",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum.per_second": "utchmma_bf16_math_ops_per_second",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed": "utchmma_bf16_pct_of_peak_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed": "pipe_tc_cycles_active_pct_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed": "pipe_tensor_hmma_cycles_active_pct_elapsed",
    "sm__sass_inst_executed_op_utcmma.sum": "utcmma_instructions",
    "sm__inst_executed_pipe_tc.sum": "pipe_tc_instructions",
    "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum": "smem_wavefronts_utcmma_a",
    "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum": "smem_wavefronts_utcmma_b_1cta",
}
UNIT = {"Tbyte": 1e12, "Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}
out = {}
for spec in sys.argv[2:]:
    name, rest = spec.split("=", 1)
    parts = rest.split(":")
    rep, extra = parts[0], dict(p.split("=") for p in parts[1:])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    d = {"kernel": v[h.index("Kernel Name")][:120], "report": rep.split("/")[-1]}
    for k, u, val in zip(h, units, v):
        if k in WANT and val != "":
            try:
                d[WANT[k]] = float(val.replace(",", "")) * UNIT.get(u, 1.0)
            except ValueError:
                pass
    for k in ("workload", "case"):
        if k in extra:
            d[k] = extra[k]
    if "flops" in extra:
        d["algorithmic_flops"] = float(extra["flops"])
        d["algorithmic_tflops_under_ncu"] = float(extra["flops"]) / (d["duration_us"] * 1e-6) / 1e12
    if "bytes" in extra:
        d["algorithmic_bytes"] = float(extra["bytes"])
        d["algorithmic_gbs_under_ncu"] = float(extra["bytes"]) / (d["duration_us"] * 1e-6) / 1e9
    out[name] = d
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
