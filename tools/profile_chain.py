"""The config-1 chain (real scan 0 -> 64x2048 image -> 284 M voxels -> mesh -> cast) a few times, for ncu:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/chain.csv python tools/profile_chain.py [reps] [sparse 0|1] [voxel]"""
import os, sys, zipfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_transfer_b200 import engine
from lidar_transfer_b200.rays import create_rays
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sparse = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
vox = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
z = zipfile.ZipFile(os.path.join(ROOT, "tests", "golden", "minimal_fixture.zip"))
scan = np.frombuffer(z.read("minimal/sequences/00/velodyne/000000.bin"), np.float32).reshape(-1, 4)
label = np.frombuffer(z.read("minimal/sequences/00/labels/000000.label"), np.uint32) & 0xFFFF
keep = ~np.isin(label, [0, 1])
pts, lab = scan[keep], label[keep]
dev = torch.device("cuda")
p64 = torch.from_numpy(pts[:, :3].astype(np.float64)).to(dev); rem = torch.from_numpy(pts[:, 3].copy()).to(dev)
lb = torch.from_numpy(lab.view(np.int32).copy()).to(dev)
bnds = np.array([[-50, 50], [-31, 40], [-3, 2]], np.float64)
dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
beams = engine.Beams(create_rays(3.0, -25.0, 64, 2048), 64)
vol = engine.TsdfDevice(dim, bnds[:, 0].astype(np.float32), vox, 3.0, -25.0, sparse=sparse)
origin = torch.zeros(3, device=dev)
for r in range(reps):
  pr = engine.project(p64, rem, lb, 3.0, -25.0, 64, 2048)
  vol.reset()
  vol.integrate(pr["proj_label"].to(torch.float32) * 65536.0, pr["range_image"], pr["proj_remissions"])
  m = vol.extract_mesh(want_norms=False, want_faces=False)
  out = engine.cast(beams, m["verts"], None, m["colors"], m["rem"], origin, zero_misses=True, check_mesh=False)
torch.cuda.synchronize()
print("tris", m["n_tris"], "hit", float((out["range"] > 0).float().mean()))
