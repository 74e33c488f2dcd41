"""How well do the latency-bound kernels of different scans overlap?  Times build-only, trace-only and
build+trace over S streams (device-resident meshes), reporting microseconds per scan."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import synth, engine, _lib, pipeline
from lidar_transfer_b200.rays import create_rays
from lidar_transfer_b200._lib import check

H, W = 64, 2048
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 710
L = _lib.lib()
scenes = [synth.make_scene(1000 + k, n_side=n_side) for k in range(4)]
dev = torch.device("cuda")
d_scenes = [tuple(torch.from_numpy(sc[k].reshape(-1)).to(dev) for k in ("verts", "faces", "colors", "rem")) for sc in scenes]
rays = create_rays(3.0, -25.0, H, W)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for S in (tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (1, 2, 4, 8, 16)):
  rr = pipeline.ScanRenderer(rays, np.zeros(3, np.float32), H, max(s["verts"].shape[0] for s in scenes),
                             max(s["faces"].shape[0] for s in scenes), n_streams=S, device=dev)
  def run(kind, reps=32):
    for i in range(reps):
      s = rr.slots[i % S]
      v, f, c, r = d_scenes[i % len(d_scenes)]
      st = ctypes.c_void_p(s.stream.cuda_stream)
      if kind in ("build", "both"):
        check(L.vl_bvh_build(P(v), P(f), P(c), P(r), v.numel() // 3, f.numel() // 3, P(s.blob), s.blob.numel(), st))
      if kind in ("trace", "both"):
        check(L.vl_trace(P(s.blob), f.numel() // 3, P(rr.rays), P(rr.origin), rr.n_rays, H, P(s.out["endpoints"]),
                         P(s.out["endcolors"]), P(s.out["range"]), P(s.out["endrem"]), P(s.out["tri_id"]), 1, st))
  run("both", S)  # every slot holds a built BVH
  torch.cuda.synchronize()
  res = {}
  for kind in ("build1", "build2", "build3", "build", "trace", "both"):
    L_stop = int(kind[5:]) if kind.startswith("build") and kind[5:] else 0
    L.vl_debug_build_stop(L_stop)
    kind = kind[:5] if kind.startswith("build") else kind
    run(kind, 2 * S); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in rr.slots: s.stream.wait_event(e0)
    reps = 64
    run(kind, reps)
    rr.fence(); e1.record(); torch.cuda.synchronize()
    res[kind + str(L_stop)] = 1e3 * e0.elapsed_time(e1) / reps
  print("streams %2d build prefix us/scan: bounds %.1f  +morton %.1f  +sort %.1f  full %.1f" % (S, res["build1"], res["build2"], res["build3"], res["build0"]))
  res = {"build": res["build0"], "trace": res["trace0"], "both": res["both0"]}
  print("streams %2d  us/scan: build %6.1f  trace %6.1f  both %6.1f  (build+trace %6.1f)" % (S, res["build"], res["trace"], res["both"], res["build"] + res["trace"]))
