"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel launch list."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s launches %3d  total %12.0f  mean %12.0f  share %5.1f%%" % (name[:60], n, t, t / n, 100 * t / tot))
