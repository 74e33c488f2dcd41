"""Host cost of submitting one scan (ScanRenderer.submit, method cast): a mesh so small that the device is idle,
one thread and two threads (ctypes releases the GIL around the C call)."""
import json, os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import synth, pipeline
from lidar_transfer_b200.rays import create_rays

H, W, fu, fd = synth.SENSORS["HDL-64E"]
rays = create_rays(fu, fd, H, W)
origin = np.zeros(3, np.float32)
sc = synth.make_scene(1, n_side=12, n_boxes=0)
mesh = tuple(torch.from_numpy(np.ascontiguousarray(sc[n]).reshape(-1)).cuda() for n in ("verts", "faces", "colors", "rem"))
res = {}
def run(R, n):
  for _ in range(n): R.submit(*mesh)
  R.wait()
for n_thr in (1, 2, 4):
  Rs = [pipeline.ScanRenderer(rays, origin, H, 1024, 1024, n_streams=8) for _ in range(n_thr)]
  for R in Rs: run(R, 50)
  torch.cuda.synchronize()
  n = 2000
  t0 = time.perf_counter()
  th = [threading.Thread(target=run, args=(R, n)) for R in Rs]
  for t in th: t.start()
  for t in th: t.join()
  torch.cuda.synchronize()
  res["threads_%d_us_per_scan" % n_thr] = round(1e6 * (time.perf_counter() - t0) / (n * n_thr), 2)
print(json.dumps(res))
