"""Per-kernel summary of an ncu --set full capture (.ncu-rep) as JSON: duration, DRAM bytes, warp instructions, active
lanes, warps active, IPC, registers, hit rates, stall cycles per issue.   python tools/ncu_extract.py file.ncu-rep [name filter]"""
import csv, io, json, re, subprocess, sys
rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def val(r, name, scale=None):
  if name not in idx: return None
  s = r[idx[name]].replace(",", "")
  try: v = float(s)
  except ValueError: return None
  u = units[idx[name]]
  if scale == "bytes":
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
  if scale == "us":
    v *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3}.get(u, 1)
  return v
out = []
for r in rows[2:]:
  name = r[idx["Kernel Name"]]
  if flt and flt not in name: continue
  k = re.search(r"(k_[a-z0-9_]+)(<[^>(]*>)?", name)
  stalls = {}
  for h in hdr:
    m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
    if m:
      v = val(r, h)
      if v is not None and v >= 0.3: stalls[m.group(1)] = round(v, 2)
  out.append({"kernel": (k.group(1) + (k.group(2) or "")) if k else name[:50],
              "grid": r[idx["Grid Size"]] if "Grid Size" in idx else val(r, "launch__grid_size"),
              "block": r[idx["Block Size"]] if "Block Size" in idx else val(r, "launch__block_size"),
              "duration_us": round(val(r, "gpu__time_duration.sum", "us") or 0, 2),
              "dram_read_bytes": val(r, "dram__bytes_read.sum", "bytes"), "dram_write_bytes": val(r, "dram__bytes_write.sum", "bytes"),
              "dram_bytes_per_launch": (val(r, "dram__bytes_read.sum", "bytes") or 0) + (val(r, "dram__bytes_write.sum", "bytes") or 0),
              "dram_pct_of_peak": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
              "warp_instructions": val(r, "smsp__inst_executed.sum"),
              "active_lanes": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
              "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
              "ipc_per_sm": val(r, "sm__inst_executed.avg.per_cycle_elapsed"),
              "registers": val(r, "launch__registers_per_thread"),
              "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"), "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"),
              "stall_cycles_per_issue": stalls})
print(json.dumps(out, indent=1))
