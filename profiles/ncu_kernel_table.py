"""Per-kernel table from an ncu report (raw page): launches, mean duration, DRAM bytes read+written per launch, achieved DRAM
GB/s, issue-slot and warp occupancy -- the "traffic" side of the roofline (what bench.py cannot measure).  argv: rep out.md title"""
import collections, csv, io, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def f(r, name):
    try: return float(r[col[name]].replace(",", ""))
    except Exception: return 0.0
def unit(name): return units[col[name]] if name in col else ""
agg = collections.OrderedDict()
for r in data:
    k = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(k, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "issue": 0.0, "warps": 0.0, "regs": 0, "l2": 0.0})
    a["n"] += 1
    t = f(r, "gpu__time_duration.sum"); tu = unit("gpu__time_duration.sum")
    a["t"] += t * ({"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tu, 1e-3))
    for key, name in (("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum"), ("l2", "lts__t_bytes.sum")):
        v = f(r, name); u = unit(name)
        a[key] += v * ({"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0))
    a["issue"] += f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"); a["warps"] += f(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["regs"] = int(f(r, "launch__registers_per_thread"))
with open(out, "w") as fo:
    fo.write("# %s\n\n(ncu --set full, --clock-control none; per-launch times under the profiler are serialised and cold-cache: compare shares and bytes, not absolutes)\n\n" % title)
    fo.write("| kernel | launches | us/launch | DRAM read MB/launch | DRAM write MB/launch | DRAM GB/s | L2 MB/launch | issue active % | warps active % | regs |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for k, a in agg.items():
        n = a["n"]
        fo.write("| %s | %d | %.1f | %.3f | %.3f | %.0f | %.2f | %.1f | %.1f | %d |\n" % (k, n, a["t"] / n, a["rd"] / n / 1e6, a["wr"] / n / 1e6,
                 (a["rd"] + a["wr"]) / max(a["t"], 1e-9) / 1e3, a["l2"] / n / 1e6, a["issue"] / n, a["warps"] / n, a["regs"]))
