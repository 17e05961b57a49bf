"""Aggregates an ncu report's source page (CUDA + SASS correlation) over all captured launches: warp-stall samples and
executed instructions per source line.  argv: report.ncu-rep out.txt [top]"""
import collections, csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 70
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
def I(x):
    try: return int(x)
    except Exception: return 0
agg = collections.defaultdict(lambda: [0, 0, ""])
fn, nk, hdr = None, 0, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fn = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": nk += 1; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if len(r) > 8 and r[0] not in ("", "Line No"):
        try: ln = int(r[0])
        except Exception: continue
        a = agg[(fn, ln)]
        a[0] += I(r[4]); a[1] += I(r[7]); a[2] = r[1][:130]
tot_s = sum(v[0] for v in agg.values()) or 1
tot_i = sum(v[1] for v in agg.values()) or 1
with open(out, "w") as f:
    f.write("%d function sections; total stall samples %d, total warp instructions %d\n" % (nk, tot_s, tot_i))
    f.write("by stall samples:\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        f.write("%7d %5.1f%%  inst %8d  %s:%d  %s\n" % (v[0], 100.0 * v[0] / tot_s, v[1], k[0], k[1], v[2].strip()))
