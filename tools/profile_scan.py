"""Short single-stream run of the per-scan hot path (LBVH build + trace on the bench scene) for ncu.

    ncu ... python tools/profile_scan.py [n_side] [n_scans]

One stream, `n_scans` distinct meshes back to back, so that every kernel of a scan appears once per
scan in launch order; nothing here is a benchmark number."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import engine, synth  # noqa: E402
from lidar_transfer_b200.rays import create_rays  # noqa: E402

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 710
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 64, 2048
rays = torch.from_numpy(create_rays(3.0, -25.0, H, W)).cuda()
origin = torch.zeros(3, device="cuda")
blob = None
out = None
for k in range(n_scans):
  sc = synth.make_scene(1000 + k, n_side=n_side)
  bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"], blob=blob)
  blob = bvh.blob
  out = engine.trace(bvh, rays, origin, H, out=out, zero_misses=True)
  torch.cuda.synchronize()
print("hit fraction", float((out["tri_id"] >= 0).float().mean()))
