"""GPU parity of the iso-surface extraction (vl_mesh_count / vl_mesh_emit) vs the oracle, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _load(dev_vol, name, arr):
  getattr(dev_vol, name).copy_(torch.from_numpy(arr))


@pytest.mark.parametrize("shape,seed", [((32, 30, 33), 0), ((70, 9, 130), 1), ((5, 4, 3), 2), ((2, 2, 2), 3),
                                        # dz % 4 == 0: the four-cubes-per-lane sweep (k_mesh_count4)
                                        ((32, 30, 32), 4), ((70, 9, 128), 5), ((5, 4, 4), 6), ((3, 41, 100), 7), ((2, 2, 8), 8)])
def test_mesh_extract_bit_exact(engine, oracle, shape, seed):
  rng = np.random.default_rng(seed)
  g = np.stack(np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij"), -1)
  c = np.array(shape) / 2.0 + 0.3
  vol = ((np.linalg.norm(g - c, axis=-1) - min(shape) * 0.37) / 3.0 + rng.normal(0, 0.05, shape)).astype(np.float32)
  vol = np.clip(vol, -1, 1)
  color_vol = (rng.choice([40, 48, 70, 259], size=shape) * 65536).astype(np.float32)
  rem_vol = rng.random(shape).astype(np.float32)
  vox, origin = 0.1, np.array([-1.5, 2.25, -3.0], np.float32)
  ref = oracle.mesh_extract(vol, color_vol, rem_vol, np.float32(vox), origin)
  dev = engine.TsdfDevice(shape, origin, vox, 3.0, -25.0)
  _load(dev, "tsdf", vol); _load(dev, "color", color_vol); _load(dev, "rem", rem_vol)
  got = dev.extract_mesh()
  assert got["faces"].shape[0] == ref["faces"].shape[0]
  if shape != (2, 2, 2):
    assert ref["faces"].shape[0] > 0
  assert np.array_equal(got["faces"].cpu().numpy(), ref["faces"])
  assert np.array_equal(got["verts"].cpu().numpy().view(np.int32), ref["verts"].view(np.int32))
  assert np.array_equal(got["colors"].cpu().numpy(), ref["colors"])
  assert np.array_equal(got["rem"].cpu().numpy().view(np.int32), ref["rem"].view(np.int32))
  n = got["norms"].cpu().numpy()
  ln = np.linalg.norm(n, axis=1)
  assert ((np.abs(ln - 1) < 1e-4) | (ln == 0)).all()


def test_empty_volume_gives_empty_mesh(engine):
  dev = engine.TsdfDevice((16, 16, 8), np.zeros(3, np.float32), 0.5, 3.0, -25.0)  # tsdf == 1 everywhere
  m = dev.extract_mesh()
  assert m["faces"].shape == (0, 3) and m["verts"].shape == (0, 3)


@pytest.mark.parametrize("shape", [(40, 123, 100), (9, 77, 7), (6, 50, 33), (4, 300, 64), (3, 3, 1), (12, 5, 2051),
                                   (3, 1, 40), (1, 30, 30)])
def test_all_four_sweeps_give_the_same_mesh(engine, vl, shape):
  """Bit-volume sweep (128-bit loads / scalar loads) vs the case-byte sweep (four cubes per lane / one): volumes with
  several 2048-cube units per plane, ragged plane ends, z rows shorter than a 32-cube word (several rows per word),
  longer than a unit, a single z layer / y row / x plane (no cube at all)."""
  rng = np.random.default_rng(11)
  g = np.stack(np.meshgrid(*[np.arange(n, dtype=np.float32) for n in shape], indexing="ij"), -1)
  r = min(17.0, 0.4 * max(min(shape), 2))
  vol = np.clip((np.linalg.norm(g - np.array(shape, np.float32) / 2, axis=-1) - r) / 4.0 + rng.normal(0, 0.2, shape), -1, 1).astype(np.float32)
  out = []
  for mode in (0, 1, 2, 3):
    vl.vl_debug_mesh_scalar(mode)
    try:
      dev = engine.TsdfDevice(shape, np.zeros(3, np.float32), 0.1, 3.0, -25.0)
      _load(dev, "tsdf", vol)
      out.append(dev.extract_mesh(want_norms=False))
    finally:
      vl.vl_debug_mesh_scalar(0)
  a = out[0]
  if min(shape) > 1:
    assert a["faces"].shape[0] > (10000 if shape[0] == 40 else 20)
  else:
    assert a["faces"].shape[0] == 0
  for b in out[1:]:
    assert a["faces"].shape[0] == b["faces"].shape[0]
    for k in ("verts", "faces", "colors", "rem"):
      assert torch.equal(a[k], b[k]), k


def test_topology_ambiguity_is_bounded_on_the_real_scan(engine):
  """Parity of the iso-surface against scikit-image's marching_cubes_lewiner is UNPINNED (scikit-image is absent from the
  reference tree and from this image, its version unpinned by the reference).  What can be said: both algorithms put the
  same vertex on every cut cube edge; they can differ in topology only in cubes with an ambiguous sign configuration.
  tools/mesh_ambiguity.py counts those on the real scan (and checks, on the way, that the device's triangle count equals
  the case-table count over a sign volume formed independently with torch): a few per cent of the active cubes."""
  import importlib.util
  import os
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  spec = importlib.util.spec_from_file_location("mesh_ambiguity", os.path.join(root, "tools", "mesh_ambiguity.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  r = mod.run(0.1)
  assert r["triangles"] > 500000 and r["active_cubes"] > 200000
  assert r["ambiguous_cube_fraction"] < 0.10 and r["beam_fraction"] < 0.15, r
  assert mod.ambiguous_cases().sum() == 128


def test_independent_torch_emitter_and_the_asymptotic_decider_on_the_real_scan(engine):
  """tools/mesh_sensitivity.py at voxel 0.1 on the fixture's scan 0: (1) a torch restatement of the emit (table lookup,
  edge interpolation, nearest-voxel colour / remission) reproduces vl_mesh.cu's mesh bit for bit; (2) every ambiguous face
  of this TSDF has a never-written outside corner and the asymptotic decider -- what Lewiner's tables follow,
  fusion_lidar.py:407 -- resolves ALL of them the way vl_mesh.cu's table does (inside corners separated), so face ambiguity
  changes no beam; (3) the triangulation of the same polygons alone does change ranges (the polygons are not planar)."""
  import os, sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
  import mesh_sensitivity
  r = mesh_sensitivity.run(0.1)
  assert r["self_check_torch_emitter_equals_vl_mesh_bit_for_bit"] is True
  assert r["ambiguous_cubes"] > 10000 and r["ambiguous_faces_with_an_untouched_outside_corner"] == r["ambiguous_faces"]
  assert r["ambiguous_cubes_the_decider_resolves_the_other_way"] == 0
  assert r["decider_margin_inside_product_over_outside_product"]["max"] < 1.0
  t = r["topology"]
  assert t["hit_mask_flips"] == 0 and t["label_flips"] == 0 and t["range_changed_at_all"] == 0
  assert r["triangulation"]["range_changed_by_more_than_1cm"] > 1000 and r["topology_worst_case"]["range_changed_by_more_than_1cm"] > 1000
  # Lewiner's interior test where it stands alone (MC33 case 4): a tunnel in a few hundred of ~half a million active cubes
  assert 500 < r["body_diagonal_only_cubes_mc33_case_4"] < 5000
  assert r["of_those_the_interior_test_joins_by_a_tunnel"] < 0.001 * r["active_cubes"]
