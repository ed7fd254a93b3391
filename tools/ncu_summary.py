"""Summarise an .ncu-rep (here, no GPU needed): python tools/ncu_summary.py file.ncu-rep [out.csv] -> key metrics of the last kernel + top stall lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
d = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
out = []
for k in hdr:
    if k in keys or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(d[k] or 0) > 0.05):
        out.append((k, d[k]))
for k, v in out:
    print("%-90s %s" % (k, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
iS, iE = h.index("# Samples"), h.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iE]
tot = sum(int(r[iS]) for r in data)
print("total samples %d, warp instructions %d" % (tot, sum(int(r[iE]) for r in data)))
for r in sorted(data, key=lambda r: -int(r[iS]))[:18]:
    print("%6.2f%% %12s  %s" % (100.*int(r[iS])/max(1, tot), r[iE], r[1].strip()[:100]))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        w = csv.writer(f)
        for k, v in out:
            w.writerow([k, v])
