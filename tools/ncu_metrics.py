"""Print selected raw metrics of an .ncu-rep (first kernel): python tools/ncu_metrics.py file.ncu-rep [substr ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
keys = sys.argv[2:] or ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
                        "lts__t_bytes.sum ", "hmma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__registers", "smsp__inst_executed.sum ",
                        "lts__t_sectors_srcunit_tex_op_read.sum ", "dram__throughput.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
                        "sm__throughput.avg.pct", "tensor_op_hmma"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
for v in rows[2:]:
    print("kernel:", v[h.index("Kernel Name")][:80])
    for k, u, val in zip(h, units, v):
        if any(x.strip() in k for x in keys) and val not in ("0", ""):
            print("  %-90s %s %s" % (k, val, u))
