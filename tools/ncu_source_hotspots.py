"""Aggregate the source page of an ncu report by source line and by device function.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_source_hotspots.py src.csv"""
import bisect, collections, csv, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
# device function start lines of drt_kernels.cuh (current tree)
starts = []
for n, line in enumerate(open(os.path.join(ROOT, "distraytracer_b200/csrc/drt_kernels.cuh")), 1):
    m = re.match(r"^(?:static )?__(?:device|global)__ .*?\b(\w+)\(", line)
    if m:
        starts.append((n, m.group(1)))
keys = [l for l, _ in starts]
cur = None
agg = collections.defaultdict(lambda: [0, 0, 0])
lines = {}
tot = [0, 0, 0]
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 9 and r[0].isdigit() and r[2] == "-":
        s, i, t, l = int(r[6]), int(r[7]), int(r[8]), int(r[0])
        k = starts[max(0, bisect.bisect_right(keys, l) - 1)][1] if cur == "drt_kernels.cuh" else cur
        a = agg[k]; a[0] += s; a[1] += i; a[2] += t
        lines[(cur, l)] = (s, i, t, r[1])
        tot[0] += s; tot[1] += i; tot[2] += t
print("total: %d samples, %.3g warp instr, %.3g thread instr, %.1f lanes/instr" % (tot[0], tot[1], tot[2], tot[2] / tot[1]))
for k, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if s * 500 > tot[0]:
        print(f"{k:28s} {100*s/tot[0]:5.1f}% samples {100*i/tot[1]:5.1f}% warp-inst  lanes {t/max(i,1):4.1f}")
print("--- top lines")
for (f, l), (s, i, t, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{f[:16]:16s}:{l:5d} {100*s/tot[0]:5.2f}% samp {100*i/tot[1]:5.2f}% inst lanes {t/max(i,1):4.1f} | {src.strip()[:105]}")
