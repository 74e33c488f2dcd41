"""How much can the re-rendered scan depend on the two things in which vl_mesh.cu may differ from the reference's
skimage.measure.marching_cubes_lewiner (fusion_lidar.py:407; absent here, unpinned)?  Both put the same vertex on every cut
cube edge; inside a cube they can differ in
  (t) how a polygon is cut into triangles (the polygons are not planar), and
  (a) which way an AMBIGUOUS face is resolved: vl_mesh.cu always separates the inside corners, Lewiner's tables follow the
      asymptotic decider (the diagonal pair with the larger product of (value - level) is connected through the saddle).
This tool MEASURES both on the real scan at config-1 size: it re-emits the iso-surface with torch from a table -- first the
product's own table (must reproduce vl_mesh.cu's vertices bit for bit: the self-check), then (t) the same polygons fanned
from their second vertex, then (a) ambiguous faces resolved by the asymptotic decider -- casts every variant with the
product's cast and compares ranges / labels / hit masks per beam with the product mesh's.  Lewiner's interior test (tunnels
between two inside corners on a body diagonal) is not modelled.   Usage: mesh_sensitivity.py [voxel=0.05]"""
import itertools, json, os, sys, zipfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from lidar_transfer_b200 import engine
from lidar_transfer_b200.rays import create_rays
import gen_mc_table as G


def case_rows(mask, connect_bits=0, rotate=0):
  """gen_mc_table's own derivation with per-face resolution (bit f of connect_bits: the inside corners of ambiguous face f
  are connected) and fan start; (mask, 0, 0) is the product's table row."""
  return G.case_triangles(mask, connect_bits, rotate)


def product_rows(mask):
  return G.case_triangles(mask)


def emit(vol_t, color_t, rem_t, cube_idx, rows_per_cube, n_rows, dim, vox, origin, level=0.0):
  """Torch restatement of k_mesh_emit for an arbitrary table: cube_idx i64[A] (linear voxel index of the cube's corner 0),
  rows_per_cube i64[A] -> row of `n_rows` (list of triangle lists).  Returns verts f32[3T,3], colors u8[3T,3], rem f32[3T]."""
  dev = vol_t.device
  mx = max(len(r) for r in n_rows)
  tab = torch.full((len(n_rows), mx, 3), -1, dtype=torch.int64)
  for i, r in enumerate(n_rows):
    for t, tri in enumerate(r):
      tab[i, t] = torch.tensor(tri)
  tab = tab.to(dev)
  cnt = torch.tensor([len(r) for r in n_rows], device=dev)
  per = cnt[rows_per_cube]
  cube_of = torch.repeat_interleave(torch.arange(cube_idx.numel(), device=dev), per)
  first = torch.cumsum(per, 0) - per
  t_in = torch.arange(cube_of.numel(), device=dev) - first[cube_of]
  e = tab[rows_per_cube[cube_of], t_in]                     # [T, 3] edge ids
  ec = torch.tensor(G.EDGES, device=dev)                     # [12, 2] corners of an edge
  ca, cb = ec[e][..., 0], ec[e][..., 1]                      # [T, 3]
  yz = dim[1] * dim[2]
  vi = cube_idx[cube_of][:, None]
  off = lambda c: (c & 1) * yz + ((c >> 1) & 1) * dim[2] + ((c >> 2) & 1)
  flat = vol_t.reshape(-1)
  va, vb = flat[vi + off(ca)], flat[vi + off(cb)]
  tt = (torch.tensor(level, dtype=torch.float32, device=dev) - va) / (vb - va)
  x, y, z = vi // yz, (vi % yz) // dim[2], vi % dim[2]
  pv = torch.stack([(x + (ca & 1)).float(), (y + ((ca >> 1) & 1)).float(), (z + ((ca >> 2) & 1)).float()], -1)   # [T, 3, 3]
  axis = torch.where((ca ^ cb) == 1, 0, torch.where((ca ^ cb) == 2, 1, 2))
  pv = pv + torch.nn.functional.one_hot(axis, 3).float() * tt[..., None]
  ix = torch.round(pv[..., 0]).long().clamp(0, dim[0] - 1)
  iy = torch.round(pv[..., 1]).long().clamp(0, dim[1] - 1)
  iz = torch.round(pv[..., 2]).long().clamp(0, dim[2] - 1)
  ni = (ix * dim[1] + iy) * dim[2] + iz
  rgb = color_t.reshape(-1)[ni]
  cb_ = torch.floor(rgb / 65536.0)
  cg_ = torch.floor((rgb - cb_ * 65536.0) / 256.0)
  cr_ = rgb - cb_ * 65536.0 - cg_ * 256.0
  colors = torch.stack([cr_, cg_, cb_], -1).floor().long().bitwise_and(255).to(torch.uint8).reshape(-1, 3)
  rem = rem_t.reshape(-1)[ni].reshape(-1)
  o = torch.tensor(origin, dtype=torch.float32, device=dev)
  verts = (pv * torch.tensor(vox, dtype=torch.float32, device=dev) + o).reshape(-1, 3)
  return verts.contiguous(), colors.contiguous(), rem.contiguous()


def run(vox=0.05):
  z = zipfile.ZipFile(os.path.join(ROOT, "tests", "golden", "minimal_fixture.zip"))
  scan = np.frombuffer(z.read("minimal/sequences/00/velodyne/000000.bin"), np.float32).reshape(-1, 4)
  label = np.frombuffer(z.read("minimal/sequences/00/labels/000000.label"), np.uint32) & 0xFFFF
  keep = ~np.isin(label, [0, 1])
  pts, lab = scan[keep], label[keep]
  bnds = np.array([[-50, 50], [-31, 40], [-3, 2]], np.float64)
  dim = [int(v) for v in np.ceil((bnds[:, 1] - bnds[:, 0]) / vox)]
  origin = bnds[:, 0].astype(np.float32)
  pr = engine.project(pts[:, :3].astype(np.float64), pts[:, 3], lab, 3.0, -25.0, 64, 2048)
  vol = engine.TsdfDevice(dim, origin, vox, 3.0, -25.0)
  vol.integrate(pr["proj_label"].to(torch.float32) * 65536.0, pr["range_image"], pr["proj_remissions"])
  m = vol.extract_mesh(want_norms=False)
  tsdf, color, rem = vol.tsdf, vol.color, vol.rem
  dev = tsdf.device
  inside = (tsdf < 0).to(torch.uint8)
  case = torch.zeros((dim[0], dim[1], dim[2]), dtype=torch.uint8, device=dev)     # cubes at the far faces stay 0
  sub = case[:dim[0] - 1, :dim[1] - 1, :dim[2] - 1]
  for c in range(8):
    cx, cy, cz = c & 1, (c >> 1) & 1, (c >> 2) & 1
    sub |= inside[cx:dim[0] - 1 + cx, cy:dim[1] - 1 + cy, cz:dim[2] - 1 + cz] << c
  flatcase = case.reshape(-1)
  cube_idx = torch.nonzero((flatcase != 0) & (flatcase != 255)).reshape(-1)          # cube order = ascending voxel index
  cases = flatcase[cube_idx].long()
  beams = engine.Beams(create_rays(3.0, -25.0, 64, 2048), 64)
  o0 = np.zeros(3, np.float32)
  base = engine.cast(beams, m["verts"], None, m["colors"], m["rem"], o0, zero_misses=True)

  def compare(verts, colors, rem_v, name):
    out = engine.cast(beams, verts, None, colors, rem_v, o0, zero_misses=True)
    h0, h1 = base["tri_id"] >= 0, out["tri_id"] >= 0
    both = h0 & h1
    d = (out["range"] - base["range"]).abs()[both]
    l0, l1 = base["endcolors"].reshape(-1, 3)[:, 2], out["endcolors"].reshape(-1, 3)[:, 2]
    n = int(h0.sum())
    return {"variant": name, "triangles": int(verts.shape[0] // 3), "beams_hit_by_the_product_mesh": n,
            "hit_mask_flips": int((h0 != h1).sum()), "label_flips": int((l0 != l1)[both].sum()),
            "range_changed_at_all": int((d > 0).sum()), "range_changed_by_more_than_1mm": int((d > 1e-3).sum()),
            "range_changed_by_more_than_1cm": int((d > 1e-2).sum()), "range_changed_by_more_than_a_voxel": int((d > vox).sum()),
            "max_range_change_m": float(d.max()) if d.numel() else 0.0,
            "fraction_of_beams_moving_more_than_1cm": float((d > 1e-2).sum()) / max(1, n),
            "fraction_of_beams_changing_label": float((l0 != l1)[both].sum()) / max(1, n)}

  res = {"tool": "tools/mesh_sensitivity.py (real scan 0 of minimal.zip, config-1 bounds, voxel %.2f, identity 64x2048 re-render)" % vox,
         "active_cubes": int(cube_idx.numel())}
  # self-check: the product's table through the torch emitter = the product's mesh, bit for bit
  rows0 = [product_rows(k) for k in range(256)]
  v0, c0, r0 = emit(tsdf, color, rem, cube_idx, cases, rows0, dim, vox, origin)
  same = bool(torch.equal(v0.view(torch.int32), m["verts"].reshape(-1, 3).view(torch.int32)) and torch.equal(c0, m["colors"].reshape(-1, 3))
              and torch.equal(r0.view(torch.int32), m["rem"].view(torch.int32)))
  res["self_check_torch_emitter_equals_vl_mesh_bit_for_bit"] = same
  # (t) the same polygons fanned from their second vertex
  rows1 = [case_rows(k, 0, 1) for k in range(256)]
  assert [len(r) for r in rows1] == [len(r) for r in rows0]
  res["triangulation"] = compare(*emit(tsdf, color, rem, cube_idx, cases, rows1, dim, vox, origin), "every polygon fanned from its second vertex instead of its first")
  # (a) ambiguous faces by the asymptotic decider
  amb_face = torch.zeros((256, 6), dtype=torch.bool)
  for k in range(256):
    ins = [(k >> c) & 1 for c in range(8)]
    for f, cs in enumerate(G.FACES):
      amb_face[k, f] = ins[cs[0]] == ins[cs[2]] and ins[cs[1]] == ins[cs[3]] and ins[cs[0]] != ins[cs[1]]
  amb_face = amb_face.to(dev)
  yz = dim[1] * dim[2]
  off = lambda c: (c & 1) * yz + ((c >> 1) & 1) * dim[2] + ((c >> 2) & 1)
  flat = tsdf.reshape(-1)
  bits = torch.zeros_like(cases)
  for f, cs in enumerate(G.FACES):
    a = [flat[cube_idx + off(c)] for c in cs]              # value - level, level = 0
    ins0 = a[0] < 0
    p_in = torch.where(ins0, a[0] * a[2], a[1] * a[3])      # product of the INSIDE diagonal pair / of the outside pair
    p_out = torch.where(ins0, a[1] * a[3], a[0] * a[2])
    bits |= (amb_face[cases, f] & (p_in > p_out)).long() << f
  # distribution of the decider's margin: product of the inside pair / product of the outside pair, per ambiguous face
  ratios = []
  for f, cs in enumerate(G.FACES):
    a = [flat[cube_idx + off(c)] for c in cs]
    ins0 = a[0] < 0
    sel = amb_face[cases, f]
    r = (torch.where(ins0, a[0] * a[2], a[1] * a[3]) / torch.where(ins0, a[1] * a[3], a[0] * a[2]))[sel]
    ratios.append(r)
    untouched = (torch.where(ins0, torch.maximum(a[1], a[3]), torch.maximum(a[0], a[2])) == 1.0)[sel]
    res.setdefault("ambiguous_faces_with_an_untouched_outside_corner", 0)
    res["ambiguous_faces_with_an_untouched_outside_corner"] += int(untouched.sum())
  ratios = torch.cat(ratios)
  res["ambiguous_faces"] = int(ratios.numel())
  res["decider_margin_inside_product_over_outside_product"] = {"max": float(ratios.max()), "p99": float(torch.quantile(ratios[:1000000], 0.99)),
                                                                "median": float(ratios.median())}
  key = cases * 64 + bits
  uniq, inv = torch.unique(key, return_inverse=True)
  rows2 = [case_rows(int(k) // 64, int(k) % 64, 0) for k in uniq.tolist()]
  n_amb = int(amb_face[cases].any(dim=1).sum())
  n_flip = int((bits != 0).sum())
  res["ambiguous_cubes"] = n_amb
  res["ambiguous_cubes_the_decider_resolves_the_other_way"] = n_flip
  res["topology"] = compare(*emit(tsdf, color, rem, cube_idx, inv, rows2, dim, vox, origin),
                            "ambiguous faces resolved by the asymptotic decider (inside corners connected where their product is the larger)")
  # Lewiner's INTERIOR test for the one configuration where it stands alone (MC33 case 4: exactly two like-signed corners, on
  # a body diagonal): the two corners are joined by a tunnel iff, at the height t* where g(t) = At Ct - Bt Dt is extremal
  # (At .. Dt: the trilinear interpolant on the four vertical edges, A / C the edges through the two corners), t* lies in
  # (0, 1), At* and Ct* still have the corners' sign and g(t*) > 0 (their regions touch in that slice).  vl_mesh.cu's table
  # always emits the two separate caps.
  diag_pairs = [(0, 7), (1, 6), (2, 5), (3, 4)]
  n_case4 = n_tunnel = 0
  for lo_c, hi_c in diag_pairs:                      # lo_c has z = 0, hi_c = lo_c ^ 7 has z = 1
    for sign in (-1.0, 1.0):                         # the pair is inside (negative) / the pair is outside (complement case)
      mask = (1 << lo_c) | (1 << hi_c)
      want = mask if sign < 0 else (255 ^ mask)
      sel = torch.nonzero(cases == want).reshape(-1)
      if sel.numel() == 0:
        continue
      ci = cube_idx[sel]
      val = lambda c: flat[ci + off(c)]
      xa, ya = lo_c & 1, (lo_c >> 1) & 1             # column of corner lo_c; hi_c sits in the opposite column
      col = lambda x, y: (val(x | (y << 1)), val(x | (y << 1) | 4))    # (bottom, top) value of the vertical edge at (x, y)
      A0, A1 = col(xa, ya); C0, C1 = col(1 - xa, 1 - ya); B0, B1 = col(xa, 1 - ya); D0, D1 = col(1 - xa, ya)
      dA, dB, dC, dD = A1 - A0, B1 - B0, C1 - C0, D1 - D0
      qa = dA * dC - dB * dD
      qb = A0 * dC + C0 * dA - B0 * dD - D0 * dB
      t = -qb / (2 * qa)
      At, Bt, Ct, Dt = A0 + dA * t, B0 + dB * t, C0 + dC * t, D0 + dD * t
      tunnel = (qa != 0) & (t > 0) & (t < 1) & (sign * At > 0) & (sign * Ct > 0) & (At * Ct - Bt * Dt > 0)
      n_case4 += int(sel.numel()); n_tunnel += int(tunnel.sum())
  res["body_diagonal_only_cubes_mc33_case_4"] = n_case4
  res["of_those_the_interior_test_joins_by_a_tunnel"] = n_tunnel
  # the opposite rule everywhere (every ambiguous face connected): the largest effect ANY face decider could have
  bits_all = torch.zeros_like(cases)
  for f in range(6):
    bits_all |= amb_face[cases, f].long() << f
  uniq3, inv3 = torch.unique(cases * 64 + bits_all, return_inverse=True)
  rows3 = [case_rows(int(k) // 64, int(k) % 64, 0) for k in uniq3.tolist()]
  res["topology_worst_case"] = compare(*emit(tsdf, color, rem, cube_idx, inv3, rows3, dim, vox, origin),
                                       "EVERY ambiguous face resolved the other way (inside corners connected): an upper bound for any face decider")
  return res


if __name__ == "__main__":
  print(json.dumps(run(float(sys.argv[1]) if len(sys.argv) > 1 else 0.05), indent=1))
