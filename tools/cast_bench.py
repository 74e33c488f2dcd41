"""Device time of the two per-scan paths on the bench workload (1.05 M triangles, 64 x 2048 beams): beam index +
scene-streaming cast vs LBVH build + traversal, single stream and 8 scans in flight; per-stage times from the
library's own CUDA events.  usage: cast_bench.py [n_side] [cells_per_row]"""
import ctypes, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import synth, engine, _lib, pipeline
from lidar_transfer_b200.rays import create_rays

L = _lib.lib()
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 710
if len(sys.argv) > 2:
  L.vl_debug_cast_cells(int(sys.argv[2]))
if os.environ.get("VL_CAST_SPLIT") is not None:
  L.vl_debug_cast_split(int(os.environ["VL_CAST_SPLIT"]))
if os.environ.get("VL_ROW_TRIM") is not None:
  L.vl_debug_cast_row_trim(int(os.environ["VL_ROW_TRIM"]))
sensor = sys.argv[3] if len(sys.argv) > 3 else "HDL-64E"
if len(sys.argv) > 4:
  L.vl_debug_cast_ctas(int(sys.argv[4]))
methods = sys.argv[5].split(",") if len(sys.argv) > 5 else ("cast", "lbvh")
STREAMS = [int(v) for v in sys.argv[6].split(",")] if len(sys.argv) > 6 else (1, 8)
if len(sys.argv) > 7:
  L.vl_debug_cast_setup_ctas(int(sys.argv[7]))
H, W, fu, fd = synth.SENSORS[sensor]
rays = create_rays(fu, fd, H, W)
origin = np.zeros(3, np.float32)
scenes = []
for k in range(8):
  sc = synth.make_scene(1000 + k, n_side=n_side)
  scenes.append(tuple(torch.from_numpy(np.ascontiguousarray(sc[n]).reshape(-1)).cuda() for n in ("verts", "faces", "colors", "rem")))
n_t = scenes[0][1].numel() // 3
max_v = max(s[0].numel() // 3 for s in scenes); max_f = max(s[1].numel() // 3 for s in scenes)

def collect():
  n = L.vl_profile_stage_count()
  ms = (ctypes.c_double * n)(); cnt = (ctypes.c_longlong * n)()
  L.vl_profile_collect(ms, cnt)
  return {L.vl_profile_stage_name(i).decode(): round(1e3 * ms[i] / cnt[i], 2) for i in range(n) if cnt[i]}

res = dict(n_tris=n_t, n_rays=H * W, sensor=sensor)
for method in methods:
  for n_streams in STREAMS:
    R = pipeline.ScanRenderer(rays, origin, H, max_v, max_f, n_streams=n_streams, method=method)
    for rep in range(3):
      for s in scenes: R.submit(*s)
    R.wait(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    a.record()
    t0 = time.perf_counter()
    for rep in range(reps):
      for s in scenes: R.submit(*s)
    host_us = 1e6 * (time.perf_counter() - t0) / (reps * len(scenes))
    R.fence(); b.record(); torch.cuda.synchronize()
    res["%s_%dstream_host_us_per_submit" % (method, n_streams)] = round(host_us, 1)
    us = 1e3 * a.elapsed_time(b) / (reps * len(scenes))
    res["%s_%dstream_us_per_scan" % (method, n_streams)] = round(us, 1)
    res["%s_%dstream_Mrays" % (method, n_streams)] = round(H * W / us, 1)
    if n_streams == 1:
      R.use_graph = False   # per-stage events need the kernel-by-kernel path
      L.vl_profile_enable(1)
      for s in scenes: R.submit(*s)
      R.wait()
      res[method + "_stages_us"] = collect()
      L.vl_profile_enable(0)
    if method == "cast" and n_streams == 1:
      info = (ctypes.c_int * 8)()
      L.vl_cast_status(ctypes.c_void_p(R.slots[0].blob.data_ptr()), ctypes.c_void_p(R.slots[0].stream.cuda_stream), info)
      res["n_active"], res["n_units"] = info[1], info[2]
      res["hit_fraction"] = float((R.slots[0].out["tri_id"] >= 0).float().mean())
# beam index build time
L.vl_profile_enable(1)
for _ in range(5): engine.Beams(rays, H)
torch.cuda.synchronize()
res["beams_us"] = collect().get("beams")
print(json.dumps(res))
