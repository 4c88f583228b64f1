"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "lts__t_bytes.sum ", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, v in zip(hdr, vals):
        if any(h.startswith(w) for w in want) and "realtime" not in h and ".max" not in h and ".min" not in h and "not_issued" not in h:
            print(f"  {h} = {v}")
