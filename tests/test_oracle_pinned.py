"""Pins the oracle (oracle/vl_oracle.c + oracle/oracle.py) against golden vectors produced by the
REFERENCE ITSELF (tests/golden/make_golden.py: the reference's Python imported from /root/reference,
its C++ ray tracer and its CUDA integrate kernel string compiled from their own sources).  CPU only."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


@pytest.fixture(scope="module")
def G():
  return np.load(GOLDEN)


def _bits(a):
  return np.ascontiguousarray(a).view(np.int32)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_create_rays(oracle, G, tag):
  fu, fd, H, W = G["rays_%s_args" % tag]
  want = G["rays_" + tag]
  assert np.array_equal(_bits(oracle.create_rays(fu, fd, int(H), int(W))), _bits(want))
  from lidar_transfer_b200.rays import create_rays  # the product's host-side ray generator
  got = create_rays(fu, fd, int(H), int(W))
  assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"] and got.shape == want.shape
  assert np.array_equal(_bits(got), _bits(want))


@pytest.mark.parametrize("tag", ["src", "tgt"])
@pytest.mark.parametrize("impl", ["c", "numpy"])
def test_projection(oracle, G, tag, impl):
  fu, fd, H, W = G["proj_%s_args" % tag]
  fn = oracle.project if impl == "c" else oracle.project_numpy
  got = fn(G["proj_%s_points_f64" % tag], G["proj_%s_rem_in" % tag], G["proj_%s_label_in" % tag], fu, fd, int(H), int(W))
  assert got["n_kept"] == int(G["proj_%s_n_kept" % tag][0])
  assert np.array_equal(got["index"], G["proj_%s_index" % tag])
  assert np.array_equal(got["proj_label"], G["proj_%s_label" % tag])
  assert np.array_equal(_bits(got["range_image"]), _bits(G["proj_%s_range" % tag]))
  assert np.array_equal(_bits(got["proj_remissions"]), _bits(G["proj_%s_rem" % tag]))
  assert np.array_equal(G["proj_%s_points_f64" % tag][got["keep"]], G["proj_%s_kept_points" % tag])
  assert (got["index"] >= 0).sum() > 0.2 * H * W


@pytest.mark.parametrize("tag,remove", [("rad", True), ("rad", False), ("deg", False)])
@pytest.mark.parametrize("impl", ["c", "numpy"])
def test_projection_with_beam_angle_snapping(oracle, tag, remove, impl):
  """laserscan.py:321-327 (pitch <- nearest entry of beam_angles), recorded from the reference's own Python by
  tests/golden/make_golden_beams.py; `deg` is the list in degrees (every pitch lands on one of two entries)."""
  B = np.load(os.path.join(os.path.dirname(GOLDEN), "golden_beams_v1.npz"))
  fu, fd, H, W = B["args"]
  fn = oracle.project if impl == "c" else oracle.project_numpy
  got = fn(B["points_f32"].astype(np.float64), B["rem"], B["label"], fu, fd, int(H), int(W), remove=remove,
           beam_angles=B["beams_" + tag].tolist())
  k = "%s_%d_" % (tag, int(remove))
  assert got["n_kept"] == int(B[k + "n_kept"][0])
  assert np.array_equal(got["index"], B[k + "index"]) and np.array_equal(got["proj_label"], B[k + "label"])
  assert np.array_equal(_bits(got["range_image"]), _bits(B[k + "range"]))
  assert np.array_equal(_bits(got["proj_remissions"]), _bits(B[k + "rem"]))
  rows = np.unique(np.nonzero(got["index"] >= 0)[0]).size
  assert rows == (16 if tag == "rad" else 2)
  # without the list the same points fill a different image: the step is not a no-op on this input
  plain = fn(B["points_f32"].astype(np.float64), B["rem"], B["label"], fu, fd, int(H), int(W), remove=remove)
  assert not np.array_equal(plain["index"], got["index"])


def test_tsdf_restatement_matches_reference_kernel(oracle, G):
  vox, fu, fd = G["tsdf_args"]
  vol = oracle.tsdf_new_volume(G["tsdf_dim"])
  for rep in (1, 2):
    st = oracle.tsdf_integrate(vol, G["tsdf_origin"], vox, G["tsdf_color_im"], G["proj_src_range"], G["proj_src_rem"], fu, fd)
    assert st["n_written"] > 500
    for k in ("tsdf", "weight", "color", "rem"):
      assert np.array_equal(_bits(vol[k]), _bits(G["tsdf_%s_%d" % (k, rep)])), (k, rep)


def _scene(G, tag):
  from lidar_transfer_b200 import synth
  seed, n_side, n_boxes, H, W = (int(v) for v in G["trace_%s_args" % tag])
  return synth.make_scene(seed, n_side=n_side, n_boxes=n_boxes), H, W


@pytest.mark.parametrize("tag", ["s", "m", "o"])
def test_trace_sse_mode_bit_exact_vs_reference(oracle, G, tag):
  sc, H, W = _scene(G, tag)
  got = oracle.trace(G["trace_%s_rays" % tag], G["trace_%s_origin" % tag], sc["verts"], sc["faces"], sc["colors"],
                     sc["rem"], H, oracle.NORMALIZE_SSE)
  assert np.array_equal(got["tri_id"], G["trace_%s_tri_id" % tag])
  for k in ("endpoints", "endcolors", "range", "endrem"):
    assert np.array_equal(_bits(got[k]), _bits(G["trace_%s_%s" % (tag, k)])), k
  assert (got["tri_id"] >= 0).mean() > 0.5


@pytest.mark.parametrize("tag", ["s", "m", "o"])
def test_trace_ieee_mode_within_north_star_tolerance(oracle, G, tag):
  """The CUDA path normalises with IEEE 1/sqrt (x86 rsqrtps is not reproducible off-CPU): ids and labels
  must agree with the reference wherever its own two builds (FMA / no FMA) agree, ranges <= 1e-4 rel."""
  sc, H, W = _scene(G, tag)
  got = oracle.trace(G["trace_%s_rays" % tag], G["trace_%s_origin" % tag], sc["verts"], sc["faces"], sc["colors"],
                     sc["rem"], H, oracle.MIN_ID_TIES)
  ref_id, ref_r = G["trace_%s_tri_id" % tag], G["trace_%s_range" % tag]
  same = got["tri_id"] == ref_id
  assert same.mean() >= 0.999, same.mean()
  hit = same & (ref_id >= 0)
  rel = np.abs(got["range"][hit] - ref_r[hit]) / ref_r[hit]
  assert rel.max() <= 1e-4, rel.max()
  # where the winning triangle differs, it is an edge/vertex tie: same range within tolerance
  diff = ~same & (ref_id >= 0) & (got["tri_id"] >= 0)
  if diff.any():
    assert (np.abs(got["range"][diff] - ref_r[diff]) / ref_r[diff]).max() <= 1e-4


def test_known_answer_triangle(oracle, G):
  """auxiliary/raytracing.py:230-263: t = 47.08144x (FMA and non-FMA builds differ by 1 ulp)."""
  got = oracle.trace(np.array([[-39.5, -25.5, -1.7]], np.float32), np.zeros(3, np.float32), G["kat_verts"],
                     np.array([[0, 1, 2]], np.int32), np.array([[0, 0, 40]] * 3, np.int32),
                     np.array([.1, .2, .3], np.float32), 1, oracle.NORMALIZE_SSE)
  assert got["tri_id"][0] == 0
  assert np.array_equal(_bits(got["range"]), _bits(G["kat_range"]))
  assert abs(float(got["range"][0]) - 47.081444) < 1e-4
  assert np.allclose(got["endpoints"], [-39.52919, -25.518845, -1.7012562], rtol=1e-6)
  assert got["endcolors"].tolist() == [0, 0, 40]
  assert np.float32(got["endrem"][0]) == np.float32((np.float32(.1) + np.float32(.2) + np.float32(.3)) / np.float32(3))


def test_bruteforce_oracle_agrees_with_bvh_oracle(oracle, G):
  sc, H, W = _scene(G, "s")
  a = oracle.trace(G["trace_s_rays"], G["trace_s_origin"], sc["verts"], sc["faces"], sc["colors"], sc["rem"], H,
                   oracle.MIN_ID_TIES)
  b = oracle.trace(G["trace_s_rays"], G["trace_s_origin"], sc["verts"], sc["faces"], sc["colors"], sc["rem"], H,
                   oracle.BRUTE_FORCE)
  for k in ("tri_id", "range", "endpoints", "endcolors", "endrem"):
    assert np.array_equal(_bits(a[k]), _bits(b[k])), k


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference checkout not present")
def test_oracle_vs_live_reference_build(oracle):
  """Where the reference checkout exists the compiled reference is also checked live, on a larger mesh."""
  from lidar_transfer_b200 import synth
  sc = synth.make_scene(77, n_side=150)
  rays = oracle.create_rays(3.0, -25.0, 32, 512)
  o = np.zeros(3, np.float32)
  ref = oracle.ref_ctrace(rays, o, sc["verts"], sc["faces"], sc["colors"], sc["rem"], 32, ids=True)
  got = oracle.trace(rays, o, sc["verts"], sc["faces"], sc["colors"], sc["rem"], 32, oracle.NORMALIZE_SSE)
  for k in ("tri_id", "range", "endpoints", "endcolors", "endrem"):
    assert np.array_equal(_bits(got[k]), _bits(ref[k])), k


def test_compare_restatement_matches_reference_golden(oracle):
  """compare() + iouEval (auxiliary/laserscan.py:1181-1301, np_ioueval.py): the oracle's restatement against vectors
  produced by the reference's own functions (tests/golden/make_golden_compare.py)."""
  G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_compare_v1.npz"))
  for tag in "abc":
    a = {k: G["cmp_%s_%s" % (tag, k)] for k in ("source_color", "target_color", "source_label", "target_label",
                                                "source_range", "target_range", "source_rem", "target_rem")}
    r = oracle.compare_numpy(nclasses=int(G["cmp_%s_nclasses" % tag]), **a)
    for k in ("label_diff", "range_diff", "rem_diff"):
      assert np.array_equal(r[k], G["cmp_%s_%s" % (tag, k)]), (tag, k)
    m_iou, m_acc, mse = G["cmp_%s_scalars" % tag]
    assert r["m_iou"] == m_iou and r["m_acc"] == m_acc and np.float64(r["mse"]) == mse, (tag, r["m_iou"], r["m_acc"], r["mse"])
    assert r["conf"].sum() == a["source_label"].size


@pytest.mark.parametrize("method", ["pdist", "depthfast"])
@pytest.mark.parametrize("remove", [1, 0])
def test_projection_methods_pdist_and_depthfast_match_the_reference(oracle, method, remove):
  """The two projection methods no caller of the reference selects (laserscan.py:392-437), restated in
  oracle.project_numpy and pinned to the reference's own Python (tests/golden/make_golden_methods.py)."""
  M = np.load(os.path.join(os.path.dirname(GOLDEN), "golden_methods_v1.npz"))
  fu, fd, H, W = M["args"]
  o = oracle.project_numpy(M["points_f32"].astype(np.float64), M["rem"], M["label"], fu, fd, int(H), int(W), remove=bool(remove),
                           method=method)
  k = "%s_%d_" % (method, remove)
  assert o["n_kept"] == int(M[k + "n_kept"][0])
  assert np.array_equal(o["index"], M[k + "index"]) and (o["index"] >= 0).sum() > 1500
  assert np.array_equal(o["range_image"].view(np.int32), M[k + "range"].view(np.int32))
  assert np.array_equal(o["proj_remissions"], M[k + "rem"])
  if method == "pdist":
    assert np.array_equal(o["proj_label"], M[k + "label"].astype(np.int32)) and (M[k + "rem"] == -1).all()
  else:
    assert (M[k + "range"][M[k + "index"] < 0] == -1).all()
