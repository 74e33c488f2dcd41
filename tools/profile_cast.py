"""Short single-stream run of the scene-streaming cast on the bench scene, for ncu.

    ncu ... python tools/profile_cast.py [n_side] [n_scans] [cells_per_row]

One stream, `n_scans` distinct meshes back to back; nothing here is a benchmark number."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import engine, synth, _lib  # noqa: E402
from lidar_transfer_b200.rays import create_rays  # noqa: E402

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 710
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if len(sys.argv) > 3:
  _lib.lib().vl_debug_cast_cells(int(sys.argv[3]))
H, W = 64, 2048
rays = torch.from_numpy(create_rays(3.0, -25.0, H, W)).cuda()
origin = torch.zeros(3, device="cuda")
beams = engine.Beams(rays, H)
ws = None
out = None
for k in range(n_scans):
  sc = synth.make_scene(1000 + k, n_side=n_side)
  if ws is None:
    ws = beams.workspace(sc["faces"].shape[0] + 4096)
  out = engine.cast(beams, sc["verts"], sc["faces"], sc["colors"], sc["rem"], origin, out=out, zero_misses=True, workspace=ws)
  torch.cuda.synchronize()
print("hit fraction", float((out["tri_id"] >= 0).float().mean()))
