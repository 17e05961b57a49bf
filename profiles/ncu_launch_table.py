"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total ms, share, mean us.
argv: launches.csv out.md "command line that was profiled" """
import collections, csv, sys
src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    v = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[col["Metric Unit"]], 1e-3)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values()) or 1.0
with open(out, "w") as f:
    f.write("# ncu launch list: `%s`\n\n(`--metrics gpu__time_duration.sum --clock-control none`; per-launch times under the profiler are serialised and cold-cache: compare SHARES, not absolutes)\n\n" % cmd)
    f.write("%d launches, %.1f ms of kernel time\n\n| kernel | launches | total ms | share | mean us |\n|---|---:|---:|---:|---:|\n" % (sum(a[0] for a in agg.values()), tot / 1e3))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| %s | %d | %.1f | %.1f %% | %.1f |\n" % (k, a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / a[0]))
