"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py file.ncu-rep [--source KERNEL_ID N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.avg.per_cycle_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_uniform.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [(h, hdr.index(h)) for h in want if h in hdr]
for r in rows[2:]:
    print("----", r[hdr.index("ID")])
    for h, i in idx:
        print(f"  {h:75s} {r[i]:>18s} {units[i]}")
    # stall reasons
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.3:
                print(f"  STALL {h[len('smsp__average_'):]:68s} {v:10.2f}")
