"""SASS opcode histogram per kernel of lidar_transfer_b200/libvlidar.so (cuobjdump -sass, no GPU needed): which
instruction classes each kernel is made of, and whether the Blackwell / Hopper async machinery appears --
UBLKCP / UTMALDG / UTMASTG (TMA: cp.async.bulk[.tensor]), SYNCS (mbarrier), UTC*MMA / LDTM / STTM (tcgen05), HMMA
(legacy tensor path), LDGSTS (cp.async), RED / ATOM (atomics).
    python tools/sass_histogram.py > profiles/r02_sass_opcodes.json"""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "lidar_transfer_b200", "libvlidar.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kernels, cur, arch = collections.OrderedDict(), None, None
for line in txt.splitlines():
  m = re.search(r"arch = (sm_\w+)", line)
  if m:
    arch = m.group(1)
  m = re.search(r"Function : (\S+)", line)
  if m:
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    k = re.search(r"(k_[a-z0-9_]+)(<[^>(]*>)?", name)
    cur = (k.group(1) + (k.group(2) or "")) if k else name[:60]
    kernels.setdefault(cur, collections.Counter())
    continue
  m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
  if m and cur:
    kernels[cur][m.group(1).split(".")[0]] += 1
special = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "HMMA", "LDGSTS", "RED", "ATOM", "ATOMG", "ATOMS", "MUFU", "REDUX", "SHFL", "BAR", "FFMA", "FMUL", "FADD", "LDG", "STG", "LDS", "STS", "DFMA")
out = {"library": "lidar_transfer_b200/libvlidar.so", "arch": arch, "kernels": {}}
for k, c in kernels.items():
  tot = sum(c.values())
  out["kernels"][k] = {"instructions": tot, "top": dict(c.most_common(12)), "special": {s: c[s] for s in special if c.get(s)}}
out["summary"] = {
    "kernels_with_TMA_bulk_copy(UBLKCP)": [k for k, c in kernels.items() if c.get("UBLKCP")],
    "kernels_with_mbarrier(SYNCS)": [k for k, c in kernels.items() if c.get("SYNCS")],
    "kernels_with_tcgen05(UTC*MMA/LDTM/STTM)": [k for k, c in kernels.items() if any(c.get(s) for s in ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM"))],
    "kernels_with_HMMA": [k for k, c in kernels.items() if c.get("HMMA")],
    "note": "no dense contraction exists on this path (BASELINE.json north_star: 'no tensor cores'); the one bulk copy is the staged top of the LBVH in k_trace_persistent",
}
print(json.dumps(out, indent=1))
