"""Summarise an ncu report (raw page CSV) into the handful of counters profiles/ keeps."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("%-75s %-12s %s" % (h, u, v))

# machine-readable DRAM traffic of the launch, for bench.py's roofline.traffic
if len(sys.argv) > 2:
    import json
    d = dict(zip(hdr, vals)); un = dict(zip(hdr, units))
    def to_bytes(k):
        v = float(d[k]); u = un[k].lower()
        return int(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u])
    json.dump({"workload": "config3: QVRCNN QP=32, 64x1920x1080 synthetic luma frames per GPU", "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
               "dram_bytes_write": to_bytes("dram__bytes_write.sum"), "kernel": d["Kernel Name"], "source": sys.argv[3] if len(sys.argv) > 3 else rep},
              open(sys.argv[2], "w"), indent=1)
