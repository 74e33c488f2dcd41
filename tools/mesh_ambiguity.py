"""How far can the iso-surface of vl_mesh.cu (classic 256-case marching cubes, this project's own table) be from the
reference's skimage.measure.marching_cubes_lewiner (fusion_lidar.py:407; scikit-image is absent from the reference tree and
from this image, its version unpinned)?  Both place a vertex on every cube edge whose end values straddle the level, by the
same linear interpolation; they can only differ INSIDE a cube: (a) in cubes whose sign configuration is ambiguous -- a face
with diagonally opposite corners inside (Lewiner decides those by the asymptotic decider on the face's bilinear
interpolant, vl_mesh.cu always separates the inside corners), or two inside corners on a body diagonal (Lewiner's
interior test) -- where the TOPOLOGY may differ; (b) elsewhere only in how the same polygon is cut into triangles.
This tool counts (a) on the real scan at config-1 size and how many beams of the identity re-render end in such a cube:
an upper bound on the beams whose hit / label can depend on the topology choice.  Usage: mesh_ambiguity.py [voxel]"""
import json, os, re, sys, zipfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_transfer_b200 import engine
from lidar_transfer_b200.rays import create_rays


def ambiguous_cases():
  """bool[256]: face-ambiguous (a face whose corners alternate inside / outside) or two inside (or two outside) corners on a body diagonal."""
  corners = [(c & 1, (c >> 1) & 1, (c >> 2) & 1) for c in range(8)]
  faces = []
  for axis in range(3):
    for side in range(2):
      cs = [c for c in range(8) if corners[c][axis] == side]
      u, v = [a for a in range(3) if a != axis]
      key = {(0, 0): 0, (1, 0): 1, (1, 1): 2, (0, 1): 3}
      cs.sort(key=lambda c: key[(corners[c][u], corners[c][v])])
      faces.append(cs)
  out = np.zeros(256, bool)
  for m in range(256):
    inside = [(m >> c) & 1 for c in range(8)]
    face_amb = any(inside[f[0]] == inside[f[2]] != inside[f[1]] == inside[f[3]] for f in faces)
    n_in = sum(inside)
    diag = False
    for flip in (0, 1):   # the configuration or its complement: exactly two corners, opposite on a body diagonal
      s = [c for c in range(8) if inside[c] != flip]
      if len(s) == 2 and (s[0] ^ s[1]) == 7:
        diag = True
    out[m] = face_amb or diag
  return out


def tri_counts():
  text = open(os.path.join(ROOT, "lidar_transfer_b200", "csrc", "vl_mc_table.inc")).read()
  body = re.search(r"#define VL_MC_TRI_COUNT \{([^}]*)\}", text).group(1)
  return np.array([int(v) for v in body.split(",")], np.int64)


def run(vox=0.05):
  z = zipfile.ZipFile(os.path.join(ROOT, "tests", "golden", "minimal_fixture.zip"))
  scan = np.frombuffer(z.read("minimal/sequences/00/velodyne/000000.bin"), np.float32).reshape(-1, 4)
  label = np.frombuffer(z.read("minimal/sequences/00/labels/000000.label"), np.uint32) & 0xFFFF
  keep = ~np.isin(label, [0, 1])
  pts, lab = scan[keep], label[keep]
  bnds = np.array([[-50, 50], [-31, 40], [-3, 2]], np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  origin = bnds[:, 0].astype(np.float32)
  pr = engine.project(pts[:, :3].astype(np.float64), pts[:, 3], lab, 3.0, -25.0, 64, 2048)
  vol = engine.TsdfDevice(dim, origin, vox, 3.0, -25.0)
  vol.integrate(pr["proj_label"].to(torch.float32) * 65536.0, pr["range_image"], pr["proj_remissions"])
  m = vol.extract_mesh(want_norms=False)
  inside = (vol.tsdf < 0).to(torch.uint8)
  case = torch.zeros((dim[0] - 1, dim[1] - 1, dim[2] - 1), dtype=torch.uint8, device=inside.device)
  for c in range(8):
    cx, cy, cz = c & 1, (c >> 1) & 1, (c >> 2) & 1
    case |= inside[cx:dim[0] - 1 + cx, cy:dim[1] - 1 + cy, cz:dim[2] - 1 + cz] << c
  amb = torch.from_numpy(ambiguous_cases()).to(case.device)
  cnt = torch.from_numpy(tri_counts()).to(case.device)
  hist = torch.bincount(case.reshape(-1).long(), minlength=256)
  active = hist.clone(); active[0] = 0; active[255] = 0
  n_active, n_amb = int(active.sum()), int(active[amb].sum())
  n_tris, n_tris_amb = int((hist * cnt).sum()), int((hist * cnt)[amb].sum())
  assert n_tris == m["faces"].shape[0], (n_tris, m["faces"].shape[0])
  # beams of the identity re-render that end in an ambiguous cube
  H, W = 64, 2048
  out = engine.cast(engine.Beams(create_rays(3.0, -25.0, H, W), H), m["verts"], m["faces"], m["colors"], m["rem"], np.zeros(3, np.float32))
  tid = out["tri_id"].long()
  hit = tid >= 0
  v = m["verts"].reshape(-1, 3, 3)[tid[hit]]                       # the hit triangles
  cen = v.mean(dim=1)
  cube = torch.floor((cen - torch.from_numpy(origin).to(cen.device)) / vox).long()
  for k in range(3):
    cube[:, k].clamp_(0, dim[k] - 2)
  hit_case = case[cube[:, 0], cube[:, 1], cube[:, 2]].long()
  n_hit, n_hit_amb = int(hit.sum()), int(amb[hit_case].sum())
  return dict(voxel=vox, volume="%d x %d x %d" % tuple(dim), active_cubes=n_active, ambiguous_active_cubes=n_amb,
              ambiguous_cube_fraction=n_amb / max(1, n_active), triangles=n_tris, triangles_in_ambiguous_cubes=n_tris_amb,
              beams_hit=n_hit, beams_ending_in_an_ambiguous_cube=n_hit_amb, beam_fraction=n_hit_amb / max(1, n_hit),
              ambiguous_cases_of_256=int(ambiguous_cases().sum()))


if __name__ == "__main__":
  print(json.dumps(run(float(sys.argv[1]) if len(sys.argv) > 1 else 0.05)))
