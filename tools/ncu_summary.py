"""Per-kernel launch count / mean / total duration from an ncu --csv launch list (gpu__time_duration.sum)."""
import csv, sys, collections, json
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[1 + skip:]:
  name = r[ki].split("(")[0].split("<")[0].strip()
  v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] in ("nsecond", "ns") else v
  a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print("%-34s n=%4d mean %9.1f us  share %.3f" % (k[:34], n, t / n, t / tot))
print("total us", round(tot, 1))
