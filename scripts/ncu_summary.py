"""Print the metrics we care about from an .ncu-rep (raw page)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("== kernel", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_global_loads",
            "sass__inst_executed_shared_loads", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warp_latency_per_inst_issued.ratio"]
    for k in keys:
        if k in d: print(f"{k:75s} {d[k]}")
    st = sorted(((float(v), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v), reverse=True)
    for v, k in st[:8]:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
