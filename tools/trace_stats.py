"""Traversal work statistics and kernel time per traversal variant on the bench scene (GPU box)."""
import ctypes, json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import synth, engine, _lib
from lidar_transfer_b200.rays import create_rays

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 710
H, W = 64, 2048
sc = synth.make_scene(1000, n_side=n_side)
rays = torch.from_numpy(create_rays(3.0, -25.0, H, W)).cuda()
origin = torch.zeros(3, device="cuda")
L = _lib.lib()
bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
print(bvh.status())
ref_out = None
for mode in (1, 2, 8):
  L.vl_debug_trace_mode(mode)
  stats = torch.zeros(2 * H * W, dtype=torch.int32, device="cuda")
  L.vl_debug_trace_stats(ctypes.c_void_p(stats.data_ptr()))
  out = engine.trace(bvh, rays, origin, H)
  torch.cuda.synchronize()
  L.vl_debug_trace_stats(ctypes.c_void_p(0))
  st = stats.cpu().numpy().reshape(H, W, 2)
  if mode == 2:
    os.makedirs('gpurun_out', exist_ok=True); np.save('gpurun_out/trace_stats_mode2.npy', st.astype(np.int16)); np.save('gpurun_out/trace_range.npy', out['range'].cpu().numpy().reshape(H, W).astype(np.float16))
  nodes, tris = st[..., 0], st[..., 1]
  pc = lambda a: [float(np.percentile(a, q)) for q in (50, 90, 99, 100)]
  ts = []
  for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.trace(bvh, rays, origin, H, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  if ref_out is None:
    ref_out = {k: v.clone() for k, v in out.items()}
  same = all(torch.equal(out[k].view(torch.int32), ref_out[k].view(torch.int32)) for k in ref_out)
  print(json.dumps(dict(mode=mode, same_as_per_thread=same, ms_min=min(ts), nodes_mean=float(nodes.mean()),
                        nodes_pct=pc(nodes), tris_mean=float(tris.mean()), tris_pct=pc(tris))))
