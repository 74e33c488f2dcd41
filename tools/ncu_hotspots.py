"""Per-source-line hot spots of an ncu capture taken with --import-source on (-lineinfo build): share of the warp
instructions, share of the stall samples and average active lanes per CUDA source line, per kernel.
    python tools/ncu_hotspots.py file.ncu-rep [kernel regex] [top=14]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else "k_"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + flt],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
  if len(row) >= 2 and row[0] == "Function Name":
    cur = {"name": row[1], "hdr": None, "rows": []}
    blocks.append(cur)
  elif cur is not None and row and row[0] == "Line No":
    cur["hdr"] = row
  elif cur is not None and cur["hdr"] is not None and len(row) == len(cur["hdr"]) and row[2] == "-":   # the per-line summary rows
    cur["rows"].append(row)
for b in blocks:
  h = {n: i for i, n in enumerate(b["hdr"])}
  k = re.search(r"(k_[a-z0-9_]+)", b["name"])
  def num(r, n):
    try: return float(r[h[n]].replace(",", ""))
    except (KeyError, ValueError): return 0.0
  rows = [(r[0], r[1].strip(), num(r, "Instructions Executed"), num(r, "# Samples"), num(r, "Thread Instructions Executed")) for r in b["rows"]]
  ti, ts = sum(r[2] for r in rows) or 1.0, sum(r[3] for r in rows) or 1.0
  print("== %s: %.2f M warp-inst, %d samples" % (k.group(1) if k else b["name"][:40], ti / 1e6, ts))
  for r in sorted(rows, key=lambda r: -(r[2] / ti + r[3] / ts))[:top]:
    print("  L%-5s inst %5.1f%%  samples %5.1f%%  lanes %4.1f | %s" % (r[0], 100 * r[2] / ti, 100 * r[3] / ts, r[4] / r[2] if r[2] else 0, r[1][:150]))
