"""Per-launch key metrics + warp-stall breakdown from an ncu report (raw page).  argv: report.ncu-rep out.txt"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
with open(out, "w") as f:
    for r in data:
        f.write("== %s\n" % r[col["Kernel Name"]][:160])
        for k in keys:
            if k in col:
                f.write("   %-70s %s %s\n" % (k, r[col[k]], units[col[k]]))
        st = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try: st.append((float(r[i].replace(",", "")), h))
                except Exception: pass
        for v, h in sorted(st, reverse=True)[:8]:
            f.write("   stall %-66s %.1f\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))
