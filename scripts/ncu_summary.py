"""Print the handful of ncu metrics the roofline discussion uses from a .ncu-rep (run here, no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, vals = rows[0], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__stack_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__maximum_warps_per_active_cycle_pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_alu.sum"]
for w in want:
    if w in hdr:
        print(f"{w} = {vals[hdr.index(w)]} {rows[1][hdr.index(w)]}")
st = []
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        try: st.append((float(vals[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
        except ValueError: pass
print("stalls (warps stalled per issue-active cycle):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
