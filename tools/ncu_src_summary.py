"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump:
warp-instructions executed and stall samples per CUDA source line, top lines per kernel.
usage: ncu_src_summary.py dump.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
fn = None; hdr = None; cur = None
data = collections.OrderedDict()
for r in rows:
  if not r: continue
  if r[0] == "Function Name": fn = r[1].split("(")[0].split("::")[-1]; data.setdefault(fn, collections.OrderedDict()); continue
  if r[0] == "Line No": hdr = r; continue
  if fn is None or hdr is None or len(r) < 10: continue
  if r[0] != "":   # a CUDA source line row (aggregated)
    cur = (r[0], r[1].strip()); continue
  if cur is None: continue
  d = data[fn].setdefault(cur, [0, 0, 0])
  try:
    d[0] += int(r[hdr.index("Instructions Executed")]); d[1] += int(r[hdr.index("# Samples")]); d[2] += int(r[hdr.index("Thread Instructions Executed")])
  except ValueError:
    pass
for fn, lines in data.items():
  ti = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
  if ti == 0: continue
  print("== %s: %.2f M warp-inst, %d samples" % (fn, ti / 1e6, ts))
  for (ln, src), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print("  L%-4s inst %5.1f%%  samples %5.1f%%  lanes %4.1f | %s" % (ln, 100.0 * v[0] / ti, 100.0 * v[1] / max(ts, 1), v[2] / max(v[0], 1), src[:110]))
