"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (profiles/)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("prgpu::<", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(sys.argv[2], "w") as f:
    f.write("kernel,launches,total_ms,avg_ms,share\n")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{n},{a[0]},{a[1]:.3f},{a[1] / a[0]:.4f},{a[1] / tot:.3f}\n")
print(open(sys.argv[2]).read())
