"""The reference-shaped call surface (lidar_transfer_b200/auxiliary) on the GPU: same calls, same results."""
import os

import numpy as np
import pytest

from lidar_transfer_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


@pytest.fixture(scope="module")
def G():
  return np.load(GOLDEN)


def test_c_trace_matches_reference_golden(engine, G):
  """C_Trace with the .pyx signature on the golden meshes (recorded from the reference C++ built with
  -ffp-contract=off): ctrace normalises the rays on the host with the reference's own rsqrtps + Newton step, so every
  output is the golden's bit for bit -- on a host whose rsqrtps estimate is the recording host's (same vendor); the
  tolerance asserts below are what must hold anywhere."""
  from lidar_transfer_b200.auxiliary.raytracer import RayTracerCython as rtc
  for tag in ("s", "m", "o"):
    seed, n_side, n_boxes, H, W = (int(v) for v in G["trace_%s_args" % tag])
    sc = synth.make_scene(seed, n_side=n_side, n_boxes=n_boxes)
    rays = G["trace_%s_rays" % tag].reshape(-1).copy()
    ep, ec = np.zeros(3 * H * W, np.float32), np.zeros(3 * H * W, np.int32)
    rg, rm = np.zeros(H * W, np.float32), np.zeros(H * W, np.float32)
    assert rtc.C_Trace(rays, G["trace_%s_origin" % tag].copy(), sc["verts"].reshape(-1), sc["faces"].reshape(-1),
                       sc["colors"].reshape(-1), sc["rem"], ep, ec, rg, rm, H, W) is None
    ref_r, ref_c = G["trace_%s_range" % tag], G["trace_%s_endcolors" % tag]
    hit = ref_r > 0
    assert np.array_equal(rg > 0, hit)
    assert (ec == ref_c).all(axis=None) or (ec.reshape(-1, 3) == ref_c.reshape(-1, 3)).all(axis=1).mean() > 0.999
    assert (np.abs(rg[hit] - ref_r[hit]) / ref_r[hit]).max() <= 1e-4
    assert np.abs(ep - G["trace_%s_endpoints" % tag]).max() <= 1e-4 * max(1.0, np.abs(ep).max())
    same_label = (ec.reshape(-1, 3) == ref_c.reshape(-1, 3)).all(axis=1)
    assert np.allclose(rm[same_label], G["trace_%s_endrem" % tag][same_label], rtol=0, atol=0.35)
    from oracle import oracle as O   # the checker: is this host's rsqrtps the recording host's?
    if np.array_equal(O.trace(rays, G["trace_%s_origin" % tag], sc["verts"], sc["faces"], sc["colors"], sc["rem"], H,
                              O.NORMALIZE_SSE)["range"].view(np.int32), ref_r.view(np.int32)):
      assert np.array_equal(rg.view(np.int32), ref_r.view(np.int32))
      assert np.array_equal(ec, ref_c) and np.array_equal(ep.view(np.int32), G["trace_%s_endpoints" % tag].view(np.int32))
      assert np.array_equal(rm.view(np.int32), G["trace_%s_endrem" % tag].view(np.int32))


def test_c_trace_rejects_what_cython_rejects(engine):
  from lidar_transfer_b200.auxiliary.raytracer import RayTracerCython as rtc
  sc = synth.make_scene(3, n_side=8, n_boxes=0)
  H, W = 2, 4
  f32, i32 = np.float32, np.int32
  good = dict(rays=np.ones(3 * H * W, f32), origin=np.zeros(3, f32), verts=sc["verts"].reshape(-1),
              faces=sc["faces"].reshape(-1), colors=sc["colors"].reshape(-1), rem=sc["rem"],
              ep=np.zeros(3 * H * W, f32), ec=np.zeros(3 * H * W, i32), rg=np.zeros(H * W, f32), rm=np.zeros(H * W, f32))
  call = lambda d: rtc.C_Trace(d["rays"], d["origin"], d["verts"], d["faces"], d["colors"], d["rem"], d["ep"], d["ec"],
                               d["rg"], d["rm"], H, W)
  call(good)
  for key, bad in (("rays", good["rays"].astype(np.float64)), ("faces", good["faces"].astype(np.int64)),
                   ("verts", sc["verts"]), ("ep", np.zeros(6 * H * W, f32)[::2])):
    d = dict(good); d[key] = bad
    with pytest.raises(ValueError):
      call(d)


def test_throw_rays_at_mesh_glue_matches_reference_golden(engine, G):
  """TSDFVolume.throw_rays_at_mesh on a known mesh: the 7-tuple, shapes and dtypes of fusion_lidar.py:452-455."""
  import torch
  from lidar_transfer_b200.auxiliary import fusion_lidar as fl
  seed, n_side, n_boxes, H, W = (int(v) for v in G["glue_args"])
  sc = synth.make_scene(seed, n_side=n_side, n_boxes=n_boxes)
  tv = fl.TSDFVolume.__new__(fl.TSDFVolume)
  dev = "cuda"
  tv._mesh = dict(verts=torch.from_numpy(sc["verts"]).to(dev), faces=torch.from_numpy(sc["faces"]).to(dev), norms=None,
                  colors=torch.from_numpy(sc["colors"].astype(np.uint8)).to(dev), rem=torch.from_numpy(sc["rem"]).to(dev))
  from lidar_transfer_b200.rays import create_rays
  out = tv.throw_rays_at_mesh(create_rays(3.0, -25.0, H, W), np.zeros(3, np.float32), H, W, None)
  names = ("endpoints", "ray_colors", "verts", "colors", "faces", "range_image", "rem_image")
  for name, a in zip(names, out):
    g = G["glue_" + name]
    assert a.shape == g.shape and a.dtype == g.dtype, name
  assert np.array_equal(out[1], G["glue_ray_colors"])
  hit = G["glue_range_image"] > 0
  assert np.array_equal(out[5] > 0, hit)
  assert (np.abs(out[5][hit] - G["glue_range_image"][hit]) / G["glue_range_image"][hit]).max() <= 1e-4
  assert np.array_equal(out[6].view(np.int32), G["glue_rem_image"].view(np.int32))


def test_tsdf_volume_pipeline_vs_oracle(engine, oracle):
  """deform('mergemesh') core: projection -> TSDFVolume -> integrate -> marching cubes -> ray cast, against the
  same chain of oracle restatements."""
  from lidar_transfer_b200.auxiliary import fusion_lidar as fl
  from lidar_transfer_b200.rays import create_rays
  pts, labels = synth.make_scan_points(11, 60000)
  H, W, fu, fd = 64, 1024, 3.0, -25.0
  pr = oracle.project(pts[:, :3].astype(np.float64), pts[:, 3], labels, fu, fd, H, W)
  vox = 0.2
  bnds = np.array([[-20, 20], [-20, 20], [-3, 2]], np.float64)
  proj_label3 = np.zeros((H, W, 3))
  proj_label3[:, :, 0] = pr["proj_label"]
  tv = fl.TSDFVolume(bnds.copy(), vox, fu, fd)
  tv.integrate(proj_label3, pr["range_image"], pr["proj_remissions"], np.eye(3), obs_weight=1.)
  tH, tW = 32, 512
  rays = create_rays(fu, fd, tH, tW)
  origin = np.zeros(3, np.float32)
  ep, rc, verts, colors, faces, rng_im, rem_im = tv.throw_rays_at_mesh(rays, origin, tH, tW, None)
  assert ep.shape == (tH * tW, 3) and rc.shape == (tH * tW, 3) and rng_im.shape == (tH, tW) and rem_im.shape == (tH, tW)
  assert faces.shape[0] > 20000 and verts.shape[0] == 3 * faces.shape[0] and colors.dtype == np.uint8
  # oracle chain
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  vol = oracle.tsdf_new_volume(dim)
  oracle.tsdf_integrate(vol, bnds[:, 0].astype(np.float32), vox, oracle.label_to_color_im(pr["proj_label"]),
                        pr["range_image"], pr["proj_remissions"], fu, fd)
  om = oracle.mesh_extract(vol["tsdf"], vol["color"], vol["rem"], np.float32(vox), bnds[:, 0].astype(np.float32))
  ot = oracle.trace(rays, origin, om["verts"], om["faces"], om["colors"].astype(np.int32), om["rem"], tH, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  # the two TSDF volumes may differ in <= 1e-4 of the voxels (libm ulps at pixel borders): compare images loosely
  assert abs(faces.shape[0] - om["faces"].shape[0]) <= 1e-3 * om["faces"].shape[0]
  ref_r = ot["range"].reshape(tH, tW)
  both = (ref_r > 0) & (rng_im > 0)
  assert ((ref_r > 0) != (rng_im > 0)).mean() < 1e-3
  assert both.mean() > 0.3
  close = np.abs(rng_im[both] - ref_r[both]) <= 1e-4 * ref_r[both]
  assert close.mean() > 0.999
  lab, ref_lab = rc[:, 2].reshape(tH, tW), ot["endcolors"].reshape(-1, 3)[:, 2].reshape(tH, tW)
  assert (lab[both] == ref_lab[both]).mean() > 0.999


def _lut(G):
  lut = {0: [0, 0, 0], 1: [0, 0, 255], 10: [245, 150, 100], 40: [255, 0, 255], 48: [75, 0, 75], 50: [0, 200, 255],
         70: [0, 175, 0], 72: [80, 240, 150]}
  for k in np.unique(G["scan_label"]):
    lut.setdefault(int(k), [int(k) % 251 + 1, 7, 9])
  return lut


@pytest.mark.parametrize("tag", ["src", "tgt"])
def test_semlaserscan_projection_matches_reference_golden(engine, G, tag):
  """SemLaserScan.do_range_projection_new + do_label_projection_new through vl_project vs the reference's Python."""
  from lidar_transfer_b200.auxiliary.laserscan import SemLaserScan
  fu, fd, H, W = G["proj_%s_args" % tag]
  s = SemLaserScan(int(H), int(W), 20, color_dict=_lut(G))
  s.points, s.remissions = G["proj_%s_points_f64" % tag].copy(), G["proj_%s_rem_in" % tag].copy()
  s.label = G["proj_%s_label_in" % tag].copy()
  s.colorize()
  s.do_range_projection_new(fu, fd, remove=True)
  s.do_label_projection_new()
  assert s.points.shape[0] == int(G["proj_%s_n_kept" % tag][0]) and np.array_equal(s.points, G["proj_%s_kept_points" % tag])
  assert np.array_equal(s.index, G["proj_%s_index" % tag]) and np.array_equal(s.proj_label, G["proj_%s_label" % tag])
  assert np.array_equal(s.range_image.view(np.int32), G["proj_%s_range" % tag].view(np.int32))
  assert np.array_equal(s.proj_remissions.view(np.int32), G["proj_%s_rem" % tag].view(np.int32))
  for pf in (False, True):
    s.do_reverse_projection_new(fu, fd, preserve_float=pf)
    assert np.allclose(s.back_points, G["proj_%s_back_%d" % (tag, int(pf))], rtol=0, atol=1e-9)


@pytest.mark.parametrize("tag,remove", [("rad", True), ("rad", False), ("deg", False)])
def test_semlaserscan_beam_angle_snapping_matches_reference_golden(engine, G, tag, remove):
  """SemLaserScan(beam_angles=...).do_range_projection_new through vl_project_snap vs the reference's own Python
  (tests/golden/make_golden_beams.py), including the per-pixel float row of the snapped pitch."""
  from lidar_transfer_b200.auxiliary.laserscan import SemLaserScan
  B = np.load(os.path.join(os.path.dirname(GOLDEN), "golden_beams_v1.npz"))
  fu, fd, H, W = B["args"]
  s = SemLaserScan(int(H), int(W), 20, color_dict=_lut(G), beam_angles=B["beams_" + tag].tolist())
  s.points, s.remissions, s.label = B["points_f32"].astype(np.float64), B["rem"].copy(), B["label"].copy()
  s.colorize()
  s.do_range_projection_new(fu, fd, remove=remove)
  s.do_label_projection_new()
  k = "%s_%d_" % (tag, int(remove))
  assert s.points.shape[0] == int(B[k + "n_kept"][0])
  assert np.array_equal(s.index, B[k + "index"]) and np.array_equal(s.proj_label, B[k + "label"])
  assert np.array_equal(s.range_image.view(np.int32), B[k + "range"].view(np.int32))
  assert np.array_equal(s.proj_remissions.view(np.int32), B[k + "rem"].view(np.int32))
  assert np.array_equal(s.proj_y_float, B[k + "proj_y_float"])


def test_semlaserscan_projection_with_no_surviving_point_raises_like_the_reference(engine, G):
  """Degree-valued beam_angles + remove=True: every row falls outside [0, 1]; the reference stops with an IndexError
  at laserscan.py:384 and so does the shim."""
  from lidar_transfer_b200.auxiliary.laserscan import SemLaserScan
  B = np.load(os.path.join(os.path.dirname(GOLDEN), "golden_beams_v1.npz"))
  fu, fd, H, W = B["args"]
  s = SemLaserScan(int(H), int(W), 20, color_dict=_lut(G), beam_angles=B["beams_deg"].tolist())
  s.points, s.remissions, s.label = B["points_f32"].astype(np.float64), B["rem"].copy(), B["label"].copy()
  s.colorize()
  with pytest.raises(IndexError):
    s.do_range_projection_new(fu, fd, remove=True)


def _dataset(G, tmp_path):
  scan = np.concatenate([G["scan_points_f32"], G["scan_rem"][:, None]], axis=1).astype(np.float32)
  names, labels = [], []
  for k in range(2):
    p, l = str(tmp_path / ("%06d.bin" % k)), str(tmp_path / ("%06d.label" % k))
    scan.tofile(p); G["scan_label"].astype(np.uint32).tofile(l)
    names.append(p); labels.append(l)
  return names, labels


def test_deform_cp_and_mergemesh_and_compare(engine, G, tmp_path):
  """The lidar_deform.py:393-452 sequence on a small scan: SemLaserScan reference scan, MultiSemLaserScan.deform in
  the 'cp' and 'mergemesh' adaptions, compare() for the identity re-render, write()."""
  from lidar_transfer_b200.auxiliary.laserscan import SemLaserScan, MultiSemLaserScan, compare
  names, labels = _dataset(G, tmp_path)
  lut = _lut(G)
  H, W = 64, 512
  src = dict(name="HDL-64E", beams=H, fov_up=3.0, fov_down=-25.0, fov_hor=360.0, angle_res_hor=360.0 / W)
  poses = [G["scan_pose"], G["scan_pose"]]
  scan = SemLaserScan(H, W, len(lut), lut)
  scan.open_scan(names[0], 3.0, -25.0); scan.open_label(labels[0]); scan.colorize(); scan.remove_classes([0, 1])
  scan.do_range_projection(3.0, -25.0, remove=True); scan.do_label_projection()

  # cp into the golden's target geometry: back-projected points equal the reference's
  tgt = dict(name="HDL-32E", beams=32, fov_up=10.67, fov_down=-30.67, fov_hor=360.0, angle_res_hor=360.0 / 256)
  ms = MultiSemLaserScan(src, tgt, 1, len(lut), [0, 1], [252, 259], lut, transformation=None, preserve_float=True,
                         voxel_size=0.25, vol_bnds=np.array([[-30, 30], [-30, 30], [-4, 3]]))
  ms.open_multiple_scans(names, labels, poses, 0)
  assert ms.deform('cp', poses, 0) == ([], [], [])
  assert np.allclose(ms.back_points, G["proj_tgt_back_1"], rtol=0, atol=1e-9)
  assert np.array_equal(ms.proj_range.view(np.int32), G["proj_tgt_range"].view(np.int32))
  os.makedirs(tmp_path / "velodyne"); os.makedirs(tmp_path / "labels")
  ms.write(str(tmp_path), 3)
  n_valid = int((G["proj_tgt_index"].reshape(-1) > 0).sum())
  assert os.path.getsize(tmp_path / "velodyne" / "000003.bin") <= 16 * n_valid

  # mergemesh, identity sensor: the synthetic scan must resemble the source scan
  ms = MultiSemLaserScan(src, src, 1, len(lut), [0, 1], [252, 259], lut, transformation=None, preserve_float=True,
                         voxel_size=0.25, vol_bnds=np.array([[-30, 30], [-30, 30], [-4, 3]]))
  ms.open_multiple_scans(names, labels, poses, 0)
  verts, vcolors, faces = ms.deform('mergemesh', poses, 0)
  assert faces.shape[0] > 1000 and verts.shape == (3 * faces.shape[0], 3) and vcolors.shape == verts.shape
  # the mesh comes back lazily (device-resident until somebody looks): the host view is the reference's arrays
  v_host, c_host, f_host = np.asarray(verts), np.asarray(vcolors), np.asarray(faces)
  assert v_host.dtype == np.float32 and v_host.shape == verts.shape and f_host.dtype == np.int32 and len(faces) == f_host.shape[0]
  assert c_host.shape == v_host.shape and np.array_equal(f_host.reshape(-1), np.arange(3 * f_host.shape[0]))
  assert np.array_equal(verts[:5], v_host[:5]) and np.array_equal(np.asarray(vcolors[..., ::-1] * 255), c_host[..., ::-1] * 255)
  assert ms.back_points.shape == (H * W, 3) and ms.proj_range.shape == (H, W) and ms.label_image.shape == (H, W)
  assert ms.proj_color.shape == (H, W, 3) and ms.label_color.shape == (H * W, 3)
  hit = ms.proj_range > 0
  assert hit.mean() > 0.02
  assert set(np.unique(ms.label_image[hit])) <= set(int(v) % 256 for v in np.unique(G["scan_label"]))
  label_diff, range_diff, rem_diff, m_iou, m_acc, mse = compare(scan, ms)
  assert label_diff.shape == (H, W, 3) and range_diff.shape == (H, W) and np.isfinite(mse) and 0 <= m_iou <= 1 and 0 <= m_acc <= 1
  from lidar_transfer_b200.auxiliary.laserscan import compare_device   # the same metrics through vl_compare
  d_label, d_range, d_rem, d_iou, d_acc, d_mse = compare_device(scan, ms)
  assert d_iou == m_iou and d_acc == m_acc and abs(d_mse - mse) <= 1e-6 * max(1.0, mse)
  assert np.array_equal(d_range, range_diff.astype(np.float32)) and np.abs(d_label - label_diff).max() <= 1e-6
  ms.write(str(tmp_path), 4)
  out = np.fromfile(tmp_path / "velodyne" / "000004.bin", np.float32).reshape(-1, 4)
  lab = np.fromfile(tmp_path / "labels" / "000004.label", np.uint32)
  assert out.shape[0] == lab.shape[0] == int(((ms.back_points.sum(axis=1) != 0)).sum())
