"""Per-stage device times of the whole synthesis chain at BASELINE.json config-1 size (C1: one 124 668-point
scan, 64x2048 image, voxel 0.05 m over [-50,50]x[-35.5,35.5]x[-3,2] = 2000 x 1420 x 100 = 284 M voxels), from the
library's own CUDA events: projection -> TSDF init / integrate -> mesh count / scan / emit -> cast (and, with a
second argument, LBVH build -> trace beside it)."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import synth, engine, _lib
from lidar_transfer_b200.rays import create_rays

L = _lib.lib()
if os.environ.get("VL_TSDF_SHELL") == "0":   # first integration without the shell sweep, for comparison
  L.vl_debug_tsdf_shell(0)
H, W, fu, fd = 64, 2048, 3.0, -25.0
vox = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
pts, labels = synth.make_scan_points(1, 124668)
p64 = torch.from_numpy(pts[:, :3].astype(np.float64)).cuda()
rem = torch.from_numpy(pts[:, 3].copy()).cuda()
lab = torch.from_numpy(labels.view(np.int32)).cuda()
bnds = np.array([[-50, 50], [-35.5, 35.5], [-3, 2]], np.float64)
dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
rays = torch.from_numpy(create_rays(fu, fd, H, W)).cuda()
origin = torch.zeros(3, device="cuda")

def collect():
  n = L.vl_profile_stage_count()
  ms = (ctypes.c_double * n)(); cnt = (ctypes.c_longlong * n)()
  L.vl_profile_collect(ms, cnt)
  return {L.vl_profile_stage_name(i).decode(): [round(1e3 * ms[i] / cnt[i], 1), int(cnt[i])] for i in range(n) if cnt[i]}

dev = engine.TsdfDevice(dim, bnds[:, 0].astype(np.float32), vox, fu, fd)
ws = None
res = None
for rep in range(3):
  L.vl_profile_enable(1 if rep else 0)
  pr = engine.project(p64, rem, lab, fu, fd, H, W, workspace=ws); ws = pr["workspace"]
  dev.reset()
  color_im = pr["proj_label"].to(torch.float32) * 65536.0
  dev.integrate(color_im, pr["range_image"], pr["proj_remissions"])
  m = dev.extract_mesh(want_norms=False)
  if rep == 0:
    beams = engine.Beams(rays, H)
  out = engine.cast(beams, m["verts"], m["faces"], m["colors"].to(torch.int32), m["rem"], origin, zero_misses=True, check_mesh=False)
  if len(sys.argv) > 2:   # the LBVH path beside it
    bvh = engine.Bvh(m["verts"], m["faces"], m["colors"].to(torch.int32), m["rem"])
    out2 = engine.trace(bvh, rays, origin, H, zero_misses=True)
  torch.cuda.synchronize()
  if rep:
    res = collect()
n_vox = int(np.prod(dim))
info = dict(voxel=vox, dim=[int(d) for d in dim], n_vox=n_vox, n_points=int(p64.shape[0]), n_tris=int(m["faces"].shape[0]),
            hit_fraction=float((out["tri_id"] >= 0).float().mean()), stages_us_per_launch=res,
            alg_GBps={"tsdf_integrate (fused with the volume reset: 16 B written per voxel)": round(16 * n_vox / (res["tsdf_integrate"][0] * 1e-6) / 1e9, 1),
                      "mesh_count": round(4 * n_vox / (res["mesh_count"][0] * 1e-6) / 1e9, 1),
                      "mesh_emit": round((4 * n_vox + 69 * 3 * int(m["faces"].shape[0])) / (res["mesh_emit"][0] * 1e-6) / 1e9, 1)})
print(json.dumps(info))
