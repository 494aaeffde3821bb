"""Summarise the per-op table that `tools/gpu_diag.py profile` writes into gpurun_out/diag.log (first chunk only)."""
import collections
import re
import sys

args = [a for a in sys.argv[1:] if not a.startswith("-")]
path = args[0] if args else "gpurun_out/diag.log"
rows, seen, on = [], 0, False
for line in open(path):
    if "profile B=" in line:
        seen += 1
        on = seen == 1
        continue
    if not on:
        continue
    f = line.split()
    if len(f) >= 3 and f[2] == "us":
        try:
            rows.append((f[0], float(f[1]), f[3] if len(f) > 3 else ""))
        except ValueError:
            pass
agg = collections.defaultdict(float)
for n, us, _ in rows:
    key = "conv(N=%s)" % n.split("_")[2] if n.startswith("conv") else re.sub(r"(_\d+)?(_h\d+)?$", "", n)
    agg[key] += us
tot = sum(agg.values())
print("TOTAL %.1f us" % tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print("   %-22s %9.1f us %5.1f%%" % (k, v, 100 * v / tot))
if "-v" in sys.argv:
    shapes = collections.defaultdict(list)
    for n, us, tf in rows:
        if n.startswith("conv"):
            shapes[n].append((us, tf))
    for n, l in sorted(shapes.items(), key=lambda x: -sum(u for u, _ in x[1])):
        print("   %-26s x%d  %8.1f us each  %s TFLOP/s" % (n, len(l), sum(u for u, _ in l) / len(l), l[0][1]))
