"""Device time of a LATER integration into the config-1 volume (n_frames > 1: the volume already holds a scan), shell
sweep vs every voxel through the reference arithmetic.  usage: tsdf_later_bench.py"""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import synth, engine, _lib

L = _lib.lib()
H, W, fu, fd, vox = 64, 2048, 3.0, -25.0, 0.05
bnds = np.array([[-50, 50], [-35.5, 35.5], [-3, 2]], np.float64)
dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
res = {}
for mode in (1, 0):
  L.vl_debug_tsdf_shell(mode)
  dev = engine.TsdfDevice(dim, bnds[:, 0].astype(np.float32), vox, fu, fd)
  times = []
  for k in range(4):
    pts, labels = synth.make_scan_points(1 + k, 124668)
    pr = engine.project(torch.from_numpy(pts[:, :3].astype(np.float64)).cuda(), torch.from_numpy(pts[:, 3].copy()).cuda(),
                        torch.from_numpy(labels.view(np.int32)).cuda(), fu, fd, H, W)
    color_im = pr["proj_label"].to(torch.float32) * 65536.0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dev.integrate(color_im, pr["range_image"], pr["proj_remissions"]); e1.record()
    torch.cuda.synchronize()
    times.append(round(e0.elapsed_time(e1) * 1e3, 1))
  res["shell" if mode else "plain"] = dict(us_per_integration=times, n_changed=int((dev.tsdf != 1).sum()),
                                          checksum=int(dev.tsdf.view(torch.int32).to(torch.int64).sum()))
  del dev
L.vl_debug_tsdf_shell(1)
print(json.dumps(res))
