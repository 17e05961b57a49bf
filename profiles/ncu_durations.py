"""kernel name (short), grid, block, duration from `ncu --metrics gpu__time_duration.sum --csv` on stdin"""
import csv, sys
rows = [r for r in csv.reader(l for l in sys.stdin if l.startswith('"'))]
hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
for r in rows[1:]:
    if len(r) < len(hdr): continue
    print("%-22s grid %-12s block %-12s %8s %s" % (r[col["Kernel Name"]].split("<")[0].replace("void ", ""), r[col["Grid Size"]], r[col["Block Size"]], r[col["Metric Value"]], r[col["Metric Unit"]]))
