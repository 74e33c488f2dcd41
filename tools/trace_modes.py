"""LBVH traversal variants on the bench scene (1.05 M triangles, 64 x 2048 beams): the per-ray kernel in 16 x 8 tiles (mode
2, the default) against the persistent-warp / compaction / TMA-staged kernel (mode 3), one stream and 8 scans in flight.
usage: trace_modes.py [reps]"""
import ctypes, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_transfer_b200 import _lib, engine, synth
from lidar_transfer_b200.rays import create_rays
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = _lib.lib()
H, W = 64, 2048
rays = torch.from_numpy(engine.normalize_rays(create_rays(3.0, -25.0, H, W))).cuda()   # unit vectors, on the device
sc = synth.make_scene(1000, n_side=710)
bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
origin = torch.zeros(3, device='cuda')
out = {}
res = {}
for mode in (2, 3):
  L.vl_debug_trace_mode(mode)
  outs = [engine.trace(bvh, rays, origin, H, normalize='given') for _ in range(3)]
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    engine.trace(bvh, rays, origin, H, out=outs[0], normalize='given')
  e1.record(); torch.cuda.synchronize()
  res["mode%d_1stream_us" % mode] = round(1e3 * e0.elapsed_time(e1) / reps, 1)
  streams = [torch.cuda.Stream() for _ in range(8)]
  bufs = [engine.trace(bvh, rays, origin, H, normalize='given') for _ in range(8)]
  torch.cuda.synchronize()
  e0.record()
  for i in range(reps * 8):
    with torch.cuda.stream(streams[i % 8]):
      if i < 8: streams[i % 8].wait_event(e0)
      engine.trace(bvh, rays, origin, H, out=bufs[i % 8], normalize='given')
  for s in streams: torch.cuda.current_stream().wait_stream(s)
  e1.record(); torch.cuda.synchronize()
  res["mode%d_8streams_us_per_trace" % mode] = round(1e3 * e0.elapsed_time(e1) / (reps * 8), 1)
  out[mode] = {k: v.clone() for k, v in outs[0].items()}
L.vl_debug_trace_mode(0)
res["identical"] = all(torch.equal(out[2][k].view(torch.int32), out[3][k].view(torch.int32)) for k in out[2])
res["n_tris"] = int(sc["faces"].shape[0]); res["rays"] = H * W
print(json.dumps(res))
