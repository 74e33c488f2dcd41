"""Times the reference-shaped plugin call (auxiliary.raytracer.RayTracerCython.C_Trace -> extern "C" ctrace) on
PAGEABLE numpy buffers: one scan of the bench workload per call, host -> device -> host inside the call.
    python tools/ctrace_bench.py [n_side=710] [calls=20] [wire=1]   (VLIDAR_COPY_THREADS=k selects the staging pool size;
                                                                     wire=0: every array crosses PCIe as the caller holds it)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_transfer_b200 import synth  # noqa: E402
from lidar_transfer_b200.auxiliary.raytracer import RayTracerCython as rtc  # noqa: E402
from lidar_transfer_b200.rays import create_rays  # noqa: E402

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 710
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
wire = int(sys.argv[3]) if len(sys.argv) > 3 else 1
H, W = 64, 2048
rays = create_rays(3.0, -25.0, H, W).reshape(-1)
origin = np.zeros(3, np.float32)
scenes = [synth.make_scene(1000 + k, n_side=n_side) for k in range(4)]
flat = [(s["verts"].reshape(-1).copy(), s["faces"].reshape(-1).copy(), s["colors"].reshape(-1).astype(np.int32), s["rem"].copy()) for s in scenes]
R = H * W
import ctypes
from lidar_transfer_b200 import _lib
_lib.lib().vl_ctrace_wire(wire)
times, phases = [], []
for i in range(calls + 3):
  v, f, c, r = flat[i % len(flat)]
  ep, ec, rg, rm = np.zeros(3 * R, np.float32), np.zeros(3 * R, np.int32), np.zeros(R, np.float32), np.zeros(R, np.float32)
  t0 = time.perf_counter()
  rtc.C_Trace(rays, origin, v, f, c, r, ep, ec, rg, rm, H, W)
  dt = time.perf_counter() - t0
  if i >= 3:
    times.append(dt)
    ph = (ctypes.c_double * 4)()
    _lib.lib().vl_ctrace_timing(ph)
    phases.append(list(ph))
times = np.array(times)
h2d_b, d2h_b = ctypes.c_longlong(0), ctypes.c_longlong(0)
_lib.lib().vl_ctrace_traffic(ctypes.byref(h2d_b), ctypes.byref(d2h_b))
bytes_in = h2d_b.value
print(json.dumps(dict(tool="ctrace_bench", copy_threads=os.environ.get("VLIDAR_COPY_THREADS", "default"), n_tris=int(flat[0][1].size // 3),
                      rays=R, ms_median=round(float(np.median(times)) * 1e3, 3), ms_min=round(float(times.min()) * 1e3, 3),
                      mrays_per_s=round(R / float(np.median(times)) / 1e6, 1), wire=wire, h2d_mb=round(bytes_in / 1e6, 1), d2h_mb=round(d2h_b.value / 1e6, 1),
                      hit_fraction=float((rg > 0).mean()),
                      phase_ms_median=dict(zip(("rays_cache", "stage_h2d", "cast_d2h_wait", "merge"), np.round(np.median(np.array(phases), axis=0), 3).tolist())))))
