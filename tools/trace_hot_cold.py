"""Is the trace kernel slower right after a build than when it is repeated on the same BVH?  Per-kernel device
times from the library's own events (recorded around the launch, no host gap)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import synth, engine, _lib
from lidar_transfer_b200.rays import create_rays

H, W = 64, 2048
L = _lib.lib()
scenes = [synth.make_scene(1000 + k, n_side=710) for k in range(2)]
rays = torch.from_numpy(create_rays(3.0, -25.0, H, W)).cuda()
origin = torch.zeros(3, device="cuda")

def collect():
  n = L.vl_profile_stage_count()
  ms = (ctypes.c_double * n)(); cnt = (ctypes.c_longlong * n)()
  L.vl_profile_collect(ms, cnt)
  return {L.vl_profile_stage_name(i).decode(): round(1e3 * ms[i] / cnt[i], 1) for i in range(n) if cnt[i]}

blob = None; out = None
for mode in (2, 5):
  L.vl_debug_trace_mode(mode)
  for rep in range(2):
    L.vl_profile_enable(1)
    for k in range(6):
      sc = scenes[k % 2]
      bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"], blob=blob); blob = bvh.blob
      out = engine.trace(bvh, rays, origin, H, out=out, zero_misses=True)
    a = collect()
    for k in range(6):
      out = engine.trace(bvh, rays, origin, H, out=out, zero_misses=True)
    b = collect()
    L.vl_profile_enable(0)
    torch.cuda.synchronize()
    # back-to-back without profiling events, one pair of torch events around 32 launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(32):
      out = engine.trace(bvh, rays, origin, H, out=out, zero_misses=True)
    e1.record(); torch.cuda.synchronize()
    print("mode", mode, "after build:", a, "| repeated:", b, "| 32 back-to-back us each: %.1f" % (1e3 * e0.elapsed_time(e1) / 32))
