"""GPU parity of rows (iii) projection and (iv) TSDF integration vs the oracle."""
import numpy as np
import pytest

from lidar_transfer_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n,H,n_ba,remove", [(5, 4000, 16, 16, True), (6, 124668, 64, 64, True),
                                                  (7, 30000, 32, 7, False), (8, 3000, 8, 1, False)])
def test_projection_with_beam_angles_matches_oracle(engine, oracle, seed, n, H, n_ba, remove):
  """vl_project_snap vs the oracle's restatement of laserscan.py:321-327 (itself pinned to the reference's Python by
  tests/golden/golden_beams_v1.npz); a list with duplicate and equidistant entries exercises argmin's first-minimum
  rule."""
  fu, fd, W = 3.0, -25.0, 512
  pts, labels = synth.make_scan_points(seed, n)
  points = pts[:, :3].astype(np.float64)
  points[::89] = 0.0
  rng = np.random.default_rng(seed)
  ba = np.sort(np.concatenate([np.linspace(fd, fu, n_ba) / 180.0 * np.pi, rng.uniform(-0.5, 0.1, 3)])).tolist()
  ba = ba + ba[:2]  # duplicates after the first occurrence must never win
  ref = oracle.project(points, pts[:, 3], labels, fu, fd, H, W, remove=remove, beam_angles=ba)
  got = engine.project(points, pts[:, 3], labels, fu, fd, H, W, remove=remove, beam_angles=ba)
  assert int(got["n_kept"].item()) == ref["n_kept"] > 0
  assert np.array_equal(got["keep"].cpu().numpy(), ref["keep"])
  for k in ("index", "proj_label"):
    assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
  for k in ("range_image", "proj_remissions"):
    assert np.array_equal(got[k].cpu().numpy().view(np.int32), ref[k].view(np.int32)), k
  plain = engine.project(points, pts[:, 3], labels, fu, fd, H, W, remove=remove)
  assert not np.array_equal(plain["index"].cpu().numpy(), ref["index"])


@pytest.mark.parametrize("seed,n,H,W,fu,fd", [(1, 5000, 16, 128, 3.0, -25.0), (2, 124668, 64, 2048, 3.0, -25.0),
                                             (3, 124668, 64, 2048, 10.67, -30.67)])
def test_projection_matches_oracle(engine, oracle, seed, n, H, W, fu, fd):
  pts, labels = synth.make_scan_points(seed, n)
  points = pts[:, :3].astype(np.float64)
  points[::97] = 0.0  # depth == 0 points are dropped
  ref = oracle.project(points, pts[:, 3], labels, fu, fd, H, W, remove=True)
  got = engine.project(points, pts[:, 3], labels, fu, fd, H, W, remove=True)
  assert int(got["n_kept"].item()) == ref["n_kept"]
  assert np.array_equal(got["keep"].cpu().numpy(), ref["keep"])
  for k in ("index", "proj_label"):
    assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
  for k in ("range_image", "proj_remissions"):
    assert np.array_equal(got[k].cpu().numpy().view(np.int32), ref[k].view(np.int32)), k


def test_tsdf_integrate_matches_oracle(engine, oracle):
  pts, labels = synth.make_scan_points(5, 60000)
  H, W, fu, fd = 64, 1024, 3.0, -25.0
  pr = oracle.project(pts[:, :3].astype(np.float64), pts[:, 3], labels, fu, fd, H, W)
  color_im = oracle.label_to_color_im(pr["proj_label"])
  vox = 0.25
  bnds = np.array([[-20, 20], [-16, 16], [-3, 2]], np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  origin = bnds[:, 0].astype(np.float32)
  vol = oracle.tsdf_new_volume(dim)
  dev = engine.TsdfDevice(dim, origin, vox, fu, fd)
  for rep in range(2):  # second pass exercises the same-label running average
    st = oracle.tsdf_integrate(vol, origin, vox, color_im, pr["range_image"], pr["proj_remissions"], fu, fd)
    dev.integrate(color_im, pr["range_image"], pr["proj_remissions"])
  assert st["n_written"] > 1000
  n = vol["tsdf"].size
  g = {k: getattr(dev, k).cpu().numpy() for k in ("tsdf", "weight", "color", "rem")}
  # voxels whose written/unwritten state differs can only come from libm ulp differences at pixel
  # borders (atan2f/asinf/norm3df: CUDA vs glibc); tolerance: <= 1e-4 of the voxels
  differs = (g["color"] != vol["color"]) | (g["weight"] != vol["weight"])
  assert differs.sum() <= 1e-4 * n, differs.sum()
  ok = ~differs
  for k in ("tsdf", "rem"):
    err = np.abs(g[k][ok] - vol[k][ok])
    assert (err > 1e-5).sum() <= 1e-4 * n, (k, (err > 1e-5).sum())


def test_projection_edge_cases(engine, oracle):
  """No points, every point outside the vertical FOV, exact depth ties (first index wins, laserscan.py:376-382) and
  float64 depths that collapse onto one float32 value (the sequential loop's last-smaller-wins rule)."""
  H, W, fu, fd = 8, 32, 3.0, -25.0
  z = np.zeros((0, 3), np.float64)
  got = engine.project(z, np.zeros(0, np.float32), np.zeros(0, np.uint32), fu, fd, H, W)
  assert int(got["n_kept"].item()) == 0 and (got["index"].cpu().numpy() == -1).all() and (got["range_image"].cpu().numpy() == 0).all()
  assert (got["proj_remissions"].cpu().numpy() == -1).all()
  up = np.tile(np.array([[1.0, 0.0, 5.0]]), (50, 1))  # pitch ~ 79 deg: outside the FOV
  ref = oracle.project(up, np.ones(50, np.float32), np.full(50, 40, np.uint32), fu, fd, H, W)
  got = engine.project(up, np.ones(50, np.float32), np.full(50, 40, np.uint32), fu, fd, H, W)
  assert ref["n_kept"] == 0 and int(got["n_kept"].item()) == 0 and (got["index"].cpu().numpy() == -1).all()
  # one pixel, many points: exact float64 ties, sub-float32 differences, and a nearer straggler in the middle
  base = np.array([10.0, 0.1, -1.0])
  eps = np.array([0.0, 0.0, 3e-9, -3e-9, 1e-12, -1e-12, 0.0, 2e-3, -2e-3, -2e-3, 5e-10])
  pts = base[None, :] * (1.0 + eps[:, None])
  rem = np.arange(len(eps), dtype=np.float32) / 16
  lab = (40 + np.arange(len(eps))).astype(np.uint32)
  for perm_seed in range(6):
    perm = np.random.default_rng(perm_seed).permutation(len(eps))
    ref = oracle.project(pts[perm], rem[perm], lab[perm], fu, fd, H, W)
    chk = oracle.project_numpy(pts[perm], rem[perm], lab[perm], fu, fd, H, W)
    got = engine.project(pts[perm], rem[perm], lab[perm], fu, fd, H, W)
    for k in ("index", "proj_label"):
      assert np.array_equal(ref[k], chk[k]) and np.array_equal(got[k].cpu().numpy(), ref[k]), (k, perm_seed)
    for k in ("range_image", "proj_remissions"):
      assert np.array_equal(got[k].cpu().numpy().view(np.int32), ref[k].view(np.int32)), (k, perm_seed)
    assert (ref["index"] >= 0).sum() == 1


def test_tsdf_class_switch_and_empty_image(engine, oracle):
  """An all-empty range image writes nothing; a second integration with OTHER labels takes the class-switch branch
  (overwrite iff dist < weight, no weight update, fusion_lidar.py:216-227)."""
  pts, labels = synth.make_scan_points(8, 30000)
  H, W, fu, fd = 32, 512, 3.0, -25.0
  pr = oracle.project(pts[:, :3].astype(np.float64), pts[:, 3], labels, fu, fd, H, W)
  vox = 0.4
  bnds = np.array([[-16, 16], [-16, 16], [-3, 2]], np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  origin = bnds[:, 0].astype(np.float32)
  dev = engine.TsdfDevice(dim, origin, vox, fu, fd)
  zero = np.zeros((H, W), np.float32)
  dev.integrate(zero, zero, zero)
  assert (dev.tsdf.cpu().numpy() == 1).all() and (dev.weight.cpu().numpy() == 0).all() and (dev.color.cpu().numpy() == 0).all()
  vol = oracle.tsdf_new_volume(dim)
  c1 = oracle.label_to_color_im(pr["proj_label"])
  c2 = oracle.label_to_color_im(np.where(pr["proj_label"] > 0, 99, 0))
  closer = (pr["range_image"] * np.float32(0.98)).astype(np.float32)
  for color_im, depth in ((c1, pr["range_image"]), (c2, closer), (c2, closer), (c1, pr["range_image"])):
    oracle.tsdf_integrate(vol, origin, vox, color_im, depth, pr["proj_remissions"], fu, fd)
    dev.integrate(color_im, depth, pr["proj_remissions"])
  assert len(np.unique(vol["color"])) >= 3 and (vol["weight"] > 1).any()
  n = vol["tsdf"].size
  g = {k: getattr(dev, k).cpu().numpy() for k in ("tsdf", "weight", "color", "rem")}
  differs = (g["color"] != vol["color"]) | (g["weight"] != vol["weight"])
  assert differs.sum() <= 1e-4 * n, differs.sum()
  for k in ("tsdf", "rem"):
    assert (np.abs(g[k][~differs] - vol[k][~differs]) > 1e-5).sum() <= 1e-4 * n, k


@pytest.mark.parametrize("vox,bnds", [(0.4, [[-16, 16], [-16, 16], [-3, 2]]), (0.08, [[-30, 30], [-21.3, 20], [-3, 2.05]])])
def test_tsdf_column_table_gives_the_same_bits(engine, oracle, vox, bnds):
  """vl_tsdf_integrate_ws (arctangent / double-precision image column once per z column) vs vl_tsdf_integrate
  (per voxel, the reference kernel's order): all four volumes bit for bit, over three integrations including a class
  switch; the second volume has 25 M voxels (voxel index beyond 2^24, where the float decode of fusion_lidar.py:96-98
  leaves the table)."""
  import torch
  pts, labels = synth.make_scan_points(12, 60000)
  H, W, fu, fd = 64, 1024, 3.0, -25.0
  pr = oracle.project(pts[:, :3].astype(np.float64), pts[:, 3], labels, fu, fd, H, W)
  bnds = np.array(bnds, np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  origin = bnds[:, 0].astype(np.float32)
  a = engine.TsdfDevice(dim, origin, vox, fu, fd)
  b = engine.TsdfDevice(dim, origin, vox, fu, fd)
  c1 = oracle.label_to_color_im(pr["proj_label"])
  c2 = oracle.label_to_color_im(np.where(pr["proj_label"] > 0, 99, 0))
  closer = (pr["range_image"] * np.float32(0.98)).astype(np.float32)
  for color_im, depth in ((c1, pr["range_image"]), (c2, closer), (c1, pr["range_image"])):
    a.integrate(color_im, depth, pr["proj_remissions"], use_column_table=True)
    b.integrate(color_im, depth, pr["proj_remissions"], use_column_table=False)
  for k in ("tsdf", "weight", "color", "rem"):
    assert torch.equal(getattr(a, k).view(torch.int32), getattr(b, k).view(torch.int32)), k
  assert int((a.tsdf != 1).sum()) > 300


@pytest.mark.parametrize("case", ["c1", "zero_labels", "bad_depths", "origin_inside", "fine_rows", "too_fine_rows", "steep_fov",
                                  "os1", "odd_dz"])
def test_tsdf_fresh_shell_sweep_gives_the_same_bits(engine, oracle, case):
  """The first integration into a new volume (vl_tsdf_init_integrate) brackets every voxel against the range image's
  [depth, depth + trunc] shell and evaluates the reference arithmetic only where an update cannot be ruled out:
  all four volumes must equal, bit for bit, (a) the same call with the shell sweep switched off, (b) the sweep with
  one voxel per thread instead of four and (c) the per-voxel kernel on a volume reset by vl_tsdf_init.  Cases: the
  config-1 volume (284 M voxels, float index decode beyond 2^24), pixels of label 0 (their voxels IN FRONT of the
  surface are written too), NaN / inf / negative depths, a volume around the sensor (the origin voxel's NaN pitch),
  512 and 1024 image rows (rows much finer than the voxels: many voxels sit near a row boundary and take two
  candidate rows), fields of view at and beyond the arcsine series' range, dz % 4 != 0."""
  import torch
  from lidar_transfer_b200._lib import lib
  H, W, fu, fd, vox = 64, 1024, 3.0, -25.0, 0.25
  bnds = [[-20, 20], [-16.1, 16], [-3, 2]]
  n_pts, seed = 60000, 21
  if case == "c1":
    H, W, vox, bnds, n_pts = 64, 2048, 0.05, [[-50, 50], [-35.5, 35.5], [-3.5, 1.5]], 124668
  elif case == "fine_rows":
    H, W = 512, 256
  elif case == "too_fine_rows":
    H, W = 1024, 128
  elif case == "odd_dz":
    bnds, vox = [[-20, 20], [-16.1, 16], [-3, 2.1]], 0.3   # 134 x 108 x 17 voxels
  elif case == "steep_fov":
    fu, fd = 10.67, -30.67
  elif case == "os1":
    H, fu, fd = 128, 22.5, -22.5
  elif case == "origin_inside":
    bnds, vox = [[-4, 4], [-4, 4], [-2, 2]], 0.125   # origin is a voxel corner: pt == (0, 0, 0) for one voxel
  pts, labels = synth.make_scan_points(seed, n_pts)
  if case == "origin_inside":
    pts[:, :3] *= 0.08
  pr = oracle.project(pts[:, :3].astype(np.float64), pts[:, 3], labels, fu, fd, H, W)
  lab, depth = pr["proj_label"].copy(), pr["range_image"].copy()
  rng = np.random.default_rng(seed)
  if case == "zero_labels":
    lab[rng.random(lab.shape) < 0.3] = 0
  if case == "bad_depths":
    r = rng.random(depth.shape)
    depth[r < 0.02] = np.nan
    depth[(r >= 0.02) & (r < 0.04)] = np.inf
    depth[(r >= 0.04) & (r < 0.06)] = -3.0
    depth[(r >= 0.06) & (r < 0.08)] = 1e-3
  color_im = oracle.label_to_color_im(lab)
  # later integrations take the same sweep (free space is skipped only for voxels never written): another class seen
  # 2 % closer (class switch in front of the first surface), the same again (running average), the first image again
  color_2 = oracle.label_to_color_im(np.where(lab > 0, 99, 0))
  closer = (depth * np.float32(0.98)).astype(np.float32)
  seq = [(color_im, depth)] if case == "c1" else [(color_im, depth), (color_2, closer), (color_2, closer), (color_im, depth)]
  bnds = np.array(bnds, np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  origin = bnds[:, 0].astype(np.float32)
  vols = {}
  for tag, mode in (("shell", 1), ("shell1", 2), ("plain", 0)):
    lib().vl_debug_tsdf_shell(mode)
    try:
      d = engine.TsdfDevice(dim, origin, vox, fu, fd)
      for c_im, d_im in seq:
        d.integrate(c_im, d_im, pr["proj_remissions"])
      vols[tag] = [getattr(d, k).view(torch.int32).clone() for k in ("tsdf", "weight", "color", "rem")]
      del d
    finally:
      lib().vl_debug_tsdf_shell(1)
  for other in ("plain", "shell1"):
    for k, a, b in zip(("tsdf", "weight", "color", "rem"), vols["shell"], vols[other]):
      assert torch.equal(a, b), (case, other, k, int((a != b).sum()))
    del vols[other]
  if case == "odd_dz":
    assert dim[2] % 4 != 0
  d = engine.TsdfDevice(dim, origin, vox, fu, fd)
  for c_im, d_im in seq:
    d.integrate(c_im, d_im, pr["proj_remissions"], use_column_table=False)   # vl_tsdf_init + the per-voxel kernel
  n_changed = int((d.tsdf != 1).sum())
  for k, a in zip(("tsdf", "weight", "color", "rem"), vols["shell"]):
    assert torch.equal(a, getattr(d, k).view(torch.int32)), (case, k)
  assert n_changed > (50 if case == "origin_inside" else 300), n_changed
  if case == "zero_labels":
    assert int(((d.weight > 0) & (d.tsdf == 1)).sum()) > 1000   # free space in front of label-0 pixels: weight > 0, tsdf 1
  if case != "c1":
    assert int((d.weight > 1).sum()) > 50 and len(torch.unique(d.color)) >= 3   # running average and class switch happened


def test_reverse_projection_on_device_matches_reference_golden(engine):
  """vl_reverse_project (A14, laserscan.py:475-501) against the back-projected points the reference itself produced
  (tests/golden/golden_v1.npz), both coordinate modes; float64, CUDA sin / cos vs the host's libm: <= 1e-9 m."""
  import os
  G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
  from lidar_transfer_b200.auxiliary.laserscan import LaserScan
  for tag in ("src", "tgt"):
    fu, fd, H, W = G["proj_%s_args" % tag]
    H, W = int(H), int(W)
    s = LaserScan(H, W)
    s.range_image = G["proj_%s_range" % tag]
    w = G["proj_%s_kept_points" % tag][G["proj_%s_index" % tag]]
    depth = np.linalg.norm(w, 2, axis=2)
    s.proj_x_float = 0.5 * (-np.arctan2(w[..., 1], w[..., 0]) / np.pi + 1.0) * W
    s.proj_y_float = (1.0 - (np.arcsin(w[..., 2] / depth) + abs(fd / 180 * np.pi)) / (abs(fd / 180 * np.pi) + abs(fu / 180 * np.pi))) * H
    s.proj_x, s.proj_y = s._clamp(s.proj_x_float, s.proj_y_float)
    for pf in (False, True):
      s.do_reverse_projection_new(fu, fd, preserve_float=pf)
      ref = G["proj_%s_back_%d" % (tag, int(pf))]
      assert s.back_points.shape == ref.shape and np.allclose(s.back_points, ref, rtol=0, atol=1e-9), (tag, pf)
      s.do_reverse_projection_new(fu, fd, preserve_float=pf, host=True)
      assert np.allclose(s.back_points, ref, rtol=0, atol=1e-9)


@pytest.mark.parametrize("method", ["pdist", "depthfast"])
@pytest.mark.parametrize("remove", [True, False])
def test_projection_methods_pdist_and_depthfast(engine, oracle, method, remove):
  """vl_project_select against the reference's own Python (golden_methods_v1.npz: fixture scan, 16 x 128 image, ~16 points
  per pixel) and against the oracle restatement on a random cloud with exact duplicates (equal depths, equal image
  positions): range / index / label (/ remission) images bit for bit."""
  import os
  M = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_methods_v1.npz"))
  fu, fd, H, W = M["args"]
  H, W = int(H), int(W)
  k = "%s_%d_" % (method, int(remove))
  g = engine.project(M["points_f32"].astype(np.float64), M["rem"], M["label"], fu, fd, H, W, remove=remove, method=method)
  assert int(g["n_kept"].item()) == int(M[k + "n_kept"][0])
  assert np.array_equal(g["index"].cpu().numpy(), M[k + "index"])
  assert np.array_equal(g["range_image"].cpu().numpy().view(np.int32), M[k + "range"].view(np.int32))
  if method == "pdist":
    assert np.array_equal(g["proj_label"].cpu().numpy(), M[k + "label"].astype(np.int32))
  else:
    assert np.array_equal(g["proj_remissions"].cpu().numpy(), M[k + "rem"])
  # random cloud, every fifth point repeated later on (ties in depth and in image position)
  rng = np.random.default_rng(17)
  pts = rng.normal(size=(6000, 3)) * np.array([20.0, 20.0, 2.0])
  pts = np.concatenate([pts, pts[::5], np.zeros((3, 3))])
  rem = rng.random(pts.shape[0], dtype=np.float32)
  lab = rng.integers(0, 300, pts.shape[0]).astype(np.uint32)
  o = oracle.project_numpy(pts, rem, lab, 10.0, -30.0, 24, 96, remove=remove, method=method)
  g = engine.project(pts, rem, lab, 10.0, -30.0, 24, 96, remove=remove, method=method)
  assert int(g["n_kept"].item()) == o["n_kept"]
  assert np.array_equal(g["index"].cpu().numpy(), o["index"])
  assert np.array_equal(g["range_image"].cpu().numpy().view(np.int32), o["range_image"].view(np.int32))
  assert np.array_equal(g["proj_label"].cpu().numpy(), o["proj_label"])
  if method == "depthfast":
    assert np.array_equal(g["proj_remissions"].cpu().numpy(), o["proj_remissions"])


def test_shim_projection_methods_set_the_reference_attributes(engine):
  """SemLaserScan.do_range_projection_new(method=...) of the drop-in package leaves the attributes the reference's
  methods leave (laserscan.py:392-437), values from the reference's own run."""
  import os
  from lidar_transfer_b200.auxiliary import laserscan as ls
  M = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_methods_v1.npz"))
  fu, fd, H, W = M["args"]
  lut = {int(v): [int(v) % 255, 3, 7] for v in np.unique(M["label"])}
  for method in ("pdist", "depthfast"):
    s = ls.SemLaserScan(int(H), int(W), max(lut) + 1, color_dict=lut)
    s.points, s.remissions, s.label = M["points_f32"].astype(np.float64), M["rem"].copy(), M["label"].copy()
    s.colorize()
    s.do_range_projection_new(fu, fd, remove=True, method=method)
    k = method + "_1_"
    assert s.points.shape[0] == int(M[k + "n_kept"][0])
    if method == "pdist":
      assert np.array_equal(s.index, M[k + "index"]) and np.array_equal(s.range_image, M[k + "range"])
      assert np.array_equal(s.label_image[..., 0], M[k + "label"]) and (s.proj_remissions == -1).all()
      assert s.proj_range is s.range_image or np.array_equal(s.proj_range, s.range_image)
      assert np.allclose(s.dist_image, M[k + "dist"], rtol=0, atol=1e-6)
      filled = s.index >= 0
      assert np.array_equal(np.asarray(s.proj_y_float)[filled], M[k + "proj_y_float"][filled])
    else:
      assert np.array_equal(s.proj_idx, M[k + "index"]) and np.array_equal(s.proj_range, M[k + "range"])
      assert np.array_equal(s.proj_remissions, M[k + "rem"]) and np.array_equal(s.proj_xyz, M[k + "xyz"])
      assert np.array_equal(s.range_image, s.proj_range) and (s.index == -1).all()
  with pytest.raises(SystemExit):
    s.do_range_projection_new(fu, fd, method="nearest")
