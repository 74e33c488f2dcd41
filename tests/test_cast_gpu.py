"""GPU parity of row (ii-b): beam index + scene-streaming cast (vl_beams_build / vl_cast) vs the oracle, through
the C ABI, bit for bit -- on beam grids, arbitrary ray sets, hostile meshes -- and against the LBVH path
(vl_bvh_build + vl_trace) at the benchmark's full size."""
import numpy as np
import pytest

from lidar_transfer_b200 import synth

pytestmark = pytest.mark.gpu
KEYS = ("tri_id", "range", "endpoints", "endcolors", "endrem")


def _np(out):
  return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}


def _same(got, ref, keys=KEYS, what=""):
  for k in keys:
    x, y = np.asarray(got[k]).reshape(-1), np.asarray(ref[k]).reshape(-1)
    same = x.view(np.int32) == y.view(np.int32)
    if not same.all():
      i = int(np.flatnonzero(~same)[0])
      raise AssertionError("%s %s: %d of %d differ, first at %d: %r vs %r" % (what, k, (~same).sum(), same.size, i, x[i], y[i]))


def _attrs(n_verts, seed=0):
  rng = np.random.default_rng(seed)
  colors = np.zeros((n_verts, 3), np.int32)
  colors[:, 2] = rng.choice(synth.STATIC_LABELS, n_verts)
  rem = rng.random(n_verts, dtype=np.float32)
  return colors, rem


def _run(engine, oracle, verts, faces, rays, origin, H, colors=None, rem=None, lbvh=True):
  verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
  faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
  if colors is None:
    colors, rem = _attrs(verts.shape[0])
  rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 3)
  origin = np.asarray(origin, np.float32)
  ref = oracle.trace(rays, origin, verts, faces, colors, rem, H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  beams = engine.Beams(rays, H)
  got = _np(engine.cast(beams, verts, faces, colors, rem, origin))
  _same(got, ref, what="cast vs oracle")
  if lbvh:
    bvh = engine.Bvh(verts, faces, colors, rem)
    _same(_np(engine.trace(bvh, rays, origin, H)), got, what="lbvh vs cast")
  return got, ref


@pytest.mark.parametrize("sensor", ["HDL-64E", "HDL-32E", "OS1-128"])
@pytest.mark.parametrize("origin", [(0.0, 0.0, 0.0), (3.5, -2.25, 0.4)])
def test_cast_beam_grids_bit_exact_vs_oracle(engine, oracle, sensor, origin):
  H, W, fu, fd = synth.SENSORS[sensor]
  W = W // 4
  sc = synth.make_scene(300 + H, n_side=160, n_boxes=12)
  rays = oracle.create_rays(fu, fd, H, W)
  got, _ = _run(engine, oracle, sc["verts"], sc["faces"], rays, origin, H, sc["colors"], sc["rem"])
  assert (got["tri_id"] >= 0).mean() > 0.8


def test_cast_full_resolution_vs_oracle(engine, oracle):
  sc = synth.make_scene(1300, n_side=300)
  H, W = 64, 2048
  rays = oracle.create_rays(3.0, -25.0, H, W)
  got, _ = _run(engine, oracle, sc["verts"], sc["faces"], rays, np.zeros(3, np.float32), H, sc["colors"], sc["rem"])
  assert (got["tri_id"] >= 0).mean() > 0.9


def _random_dirs(rng, n):
  d = rng.normal(size=(n, 3)).astype(np.float32)
  d *= rng.uniform(0.1, 30.0, (n, 1)).astype(np.float32)   # not normalised, like any caller's ray set
  return d


@pytest.mark.parametrize("H", [1, 7, 64])
def test_cast_arbitrary_ray_sets(engine, oracle, H):
  """Rays that are no beam grid at all: random directions over the whole sphere (both poles included), ragged
  counts, zero / NaN / inf / denormal-length directions (normalise to non-finite: must miss)."""
  rng = np.random.default_rng(50 + H)
  sc = synth.make_scene(77, n_side=50, n_boxes=10)
  rays = _random_dirs(rng, 64 * 61 + 3)
  special = np.array([[0, 0, 1], [0, 0, -1], [0, 0, -5], [1, 0, 0], [-1, 0, 0], [-1, -0.0, 0], [0, 1, 0], [0, -1, 0],
                      [0, 0, 0], [np.nan, 0, -1], [np.inf, 0, -1], [1e-30, 0, -1e-30], [1e-25, 1e-25, -1e-25],
                      [3e19, 1e19, -1e19], [1e20, 1.0, -1.0], [-1, 1e-9, -0.03], [-1, -1e-9, -0.03]], np.float32)
  rays[5:5 + len(special)] = special
  got, ref = _run(engine, oracle, sc["verts"], sc["faces"], rays, np.array([0.3, 0.2, 0.1], np.float32), H, sc["colors"], sc["rem"])
  assert (got["tri_id"] >= 0).mean() > 0.3
  assert got["tri_id"][5 + 8] == -1 and got["tri_id"][5 + 9] == -1


def _hostile_mesh(rng, n, origin):
  """Triangle soup around `origin`: tiny far triangles, huge ones that contain the z axis or the sensor's
  horizontal plane, needles, zero-area ones, ones with a vertex (almost) at the origin or on the z axis."""
  o = np.asarray(origin, np.float32)
  c = o + rng.normal(size=(n, 1, 3)).astype(np.float32) * rng.choice([0.5, 3.0, 20.0, 80.0], (n, 1, 1)).astype(np.float32)
  tri = c + rng.normal(size=(n, 3, 3)).astype(np.float32) * rng.choice([1e-3, 0.05, 0.5, 5.0, 60.0], (n, 1, 1)).astype(np.float32)
  k = n // 16
  tri[0 * k:1 * k, 2] = tri[0 * k:1 * k, 1]                                  # zero area
  tri[1 * k:2 * k, 2] = tri[1 * k:2 * k, 1] + 1e-4 * (tri[1 * k:2 * k, 0] - tri[1 * k:2 * k, 1])  # needles
  tri[2 * k:3 * k, 0] = o                                                     # a vertex at the origin
  tri[3 * k:4 * k, 0, :2] = o[:2]                                             # a vertex on the z axis through the origin
  big = rng.uniform(-1, 1, (k, 3, 3)).astype(np.float32) * 200.0              # huge: often pierced by the z axis
  big[:, :, 2] = rng.choice([-30.0, -2.0, 0.0, 2.0, 30.0], (k, 1)).astype(np.float32) + rng.normal(size=(k, 3)).astype(np.float32) * 0.5
  tri[4 * k:5 * k] = o + big
  tri[5 * k:6 * k, :, 2] = o[2]                                               # in the sensor's horizontal plane (edge-on)
  tri[6 * k:7 * k, 1] = o + (tri[6 * k:7 * k, 0] - o) * np.float32(1.5)       # an edge through the origin's ray (radial edge)
  verts = tri.reshape(-1, 3)
  faces = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
  return verts, faces


@pytest.mark.parametrize("seed,origin", [(1, (0.0, 0.0, 0.0)), (2, (10.0, -4.0, 1.5)), (3, (1.0e4, 2.0e4, 50.0))])
def test_cast_hostile_meshes(engine, oracle, seed, origin):
  rng = np.random.default_rng(seed)
  verts, faces = _hostile_mesh(rng, 4096, origin)
  grid = oracle.create_rays(25.0, -25.0, 32, 256)
  got, _ = _run(engine, oracle, verts, faces, grid, origin, 32)
  assert (got["tri_id"] >= 0).mean() > 0.5
  rays = _random_dirs(rng, 4000)
  _run(engine, oracle, verts, faces, rays, origin, 8)


def test_cast_non_finite_vertices_and_bad_faces(engine, oracle):
  from lidar_transfer_b200._lib import VlidarError, VL_EBADMESH
  sc = synth.make_scene(2, n_side=30, n_boxes=2)
  verts = sc["verts"].copy()
  verts[5] = np.nan
  verts[77, 1] = np.inf
  verts[200, 2] = -np.inf
  verts[300] = 3e38
  rays = oracle.create_rays(3.0, -25.0, 16, 128)
  got, _ = _run(engine, oracle, verts, sc["faces"], rays, np.zeros(3, np.float32), 16, sc["colors"], sc["rem"], lbvh=False)
  assert (got["tri_id"] >= 0).mean() > 0.5
  faces = sc["faces"].copy()
  faces[7, 1] = sc["verts"].shape[0] + 5
  faces[11, 0] = -1
  beams = engine.Beams(rays, 16)
  with pytest.raises(VlidarError) as e:
    engine.cast(beams, sc["verts"], faces, sc["colors"], sc["rem"], np.zeros(3, np.float32), check_mesh=True)
  assert e.value.code == VL_EBADMESH and "2 face" in str(e.value)
  out = _np(engine.cast(beams, sc["verts"], faces, sc["colors"], sc["rem"], np.zeros(3, np.float32), check_mesh=False))
  assert not np.isin(out["tri_id"], [7, 11]).any() and (out["tri_id"] >= 0).mean() > 0.5


@pytest.mark.parametrize("n_faces", [0, 1, 2, 255, 256, 257, 1025])
def test_cast_tiny_meshes_ragged_rays_and_empty(engine, oracle, n_faces):
  rng = np.random.default_rng(n_faces)
  sc = synth.make_scene(9, n_side=24, n_boxes=0)
  faces = sc["faces"][rng.permutation(sc["faces"].shape[0])[:n_faces]]
  rays = oracle.create_rays(-5.0, -40.0, 7, 33)[:7 * 33 - 5]   # height 7 -> width 32, 2 rays never cast
  got, _ = _run(engine, oracle, sc["verts"], faces, rays, np.zeros(3, np.float32), 7, sc["colors"], sc["rem"])
  assert (got["tri_id"][7 * 32:] == -1).all()
  if n_faces == 0:
    assert (got["tri_id"] == -1).all() and (got["range"] == 0).all()
  if n_faces == 2:   # no rays at all / fewer rays than rows
    for n in (0, 3):
      beams = engine.Beams(rays[:n], 7)
      out = _np(engine.cast(beams, sc["verts"], faces, sc["colors"], sc["rem"], np.zeros(3, np.float32)))
      assert out["tri_id"].shape == (n,) and (out["tri_id"] == -1).all()


def test_cast_close_geometry_is_shared_by_many_ctas(engine, oracle):
  """Triangles right in front of the sensor cover most of the cell grid: their items span many chunks."""
  wall = np.array([[1.0, -30.0, -30.0], [1.0, 30.0, -30.0], [1.0, 30.0, 30.0], [1.0, -30.0, 30.0],
                   [-0.5, -20.0, -20.0], [-0.5, 20.0, -20.0], [-0.5, 0.0, 20.0]], np.float32)
  faces = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6]], np.int32)
  sc = synth.make_scene(4, n_side=40, n_boxes=3)
  verts = np.concatenate([wall, sc["verts"]])
  faces = np.concatenate([faces, sc["faces"] + len(wall)])
  colors, rem = _attrs(verts.shape[0], 3)
  rays = oracle.create_rays(22.5, -22.5, 64, 512)
  ref = oracle.trace(rays, np.zeros(3, np.float32), verts, faces, colors, rem, 64, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  beams = engine.Beams(rays, 64)
  got = _np(engine.cast(beams, verts, faces, colors, rem, np.zeros(3, np.float32), check_mesh=True))
  _same(got, ref)
  assert got["n_units"] > 300 and 3 <= got["n_active"] < faces.shape[0] and np.isin(got["tri_id"], [0, 1, 2]).mean() > 0.5


@pytest.mark.parametrize("cells", [2, 4])
def test_cast_result_independent_of_cell_grid(engine, oracle, vl, cells):
  sc = synth.make_scene(88, n_side=100, n_boxes=8)
  rays = oracle.create_rays(3.0, -25.0, 32, 512)
  origin = np.array([0.5, 0.5, 0.0], np.float32)
  base = _np(engine.cast(engine.Beams(rays, 32), sc["verts"], sc["faces"], sc["colors"], sc["rem"], origin))
  vl.vl_debug_cast_cells(cells)
  try:
    alt = _np(engine.cast(engine.Beams(rays, 32), sc["verts"], sc["faces"], sc["colors"], sc["rem"], origin))
  finally:
    vl.vl_debug_cast_cells(1)
  _same(alt, base)
  _same(base, oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], 32, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE))


def test_cast_work_unit_overflow_is_reported_and_ctrace_falls_back(engine, oracle):
  """A mesh whose triangles all touch the sensor origin makes every beam a candidate of every triangle: the work
  units exceed the workspace, vl_cast reports VL_ENOSPACE without touching any output, and the host-pointer ctrace
  answers through the LBVH path instead."""
  from lidar_transfer_b200._lib import VlidarError, VL_ENOSPACE
  rng = np.random.default_rng(5)
  n = 700
  origin = np.array([1.0, 2.0, 0.5], np.float32)
  tri = origin + rng.normal(size=(n, 3, 3)).astype(np.float32) * 5.0
  tri[:, 0] = origin
  tri[::2, 0] = origin + np.float32(1e-3) * rng.normal(size=(n // 2, 3)).astype(np.float32)   # ... or nearly so
  verts, faces = tri.reshape(-1, 3), np.arange(3 * n, dtype=np.int32).reshape(n, 3)
  colors, rem = _attrs(verts.shape[0], 9)
  H, W = 64, 2048
  rays = oracle.create_rays(3.0, -25.0, H, W)
  ref = oracle.trace(rays, origin, verts, faces, colors, rem, H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  beams = engine.Beams(rays, H)
  import torch
  out = dict(endpoints=torch.full((3 * H * W,), 7.0, device="cuda"), endcolors=torch.full((3 * H * W,), 7, dtype=torch.int32, device="cuda"),
             range=torch.full((H * W,), 7.0, device="cuda"), endrem=torch.full((H * W,), 7.0, device="cuda"))
  with pytest.raises(VlidarError) as e:
    engine.cast(beams, verts, faces, colors, rem, origin, out=out, check_mesh=True)
  assert e.value.code == VL_ENOSPACE
  assert (out["range"] == 7).all() and (out["endcolors"] == 7).all()   # nothing was written
  got = engine.ctrace_host(rays, origin, verts.reshape(-1), faces.reshape(-1), colors.reshape(-1), rem, H, want_ids=True, method="cast")
  _same(got, ref, what="ctrace fallback")
  assert (got["tri_id"] >= 0).mean() > 0.3
  # the same kind of mesh, small enough for the unit list: the cast itself answers
  k = 12
  got2 = _np(engine.cast(beams, verts[:3 * k], faces[:k], colors[:3 * k], rem[:3 * k], origin, check_mesh=True))
  _same(got2, oracle.trace(rays, origin, verts[:3 * k], faces[:k], colors[:3 * k], rem[:3 * k], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE))
  assert got2["n_units"] > 10000


def test_cast_zero_misses_and_hits_only(engine, oracle):
  sc = synth.make_scene(21, n_side=20, n_boxes=3)
  rays = oracle.create_rays(30.0, -25.0, 16, 64)   # the upper rows look at the sky
  origin = np.zeros(3, np.float32)
  import torch
  beams = engine.Beams(rays, 16)
  n = rays.shape[0]
  out = dict(endpoints=torch.full((3 * n,), 7.0, device="cuda"), endcolors=torch.full((3 * n,), 7, dtype=torch.int32, device="cuda"),
             range=torch.full((n,), 7.0, device="cuda"), endrem=torch.full((n,), 7.0, device="cuda"))
  got = _np(engine.cast(beams, sc["verts"], sc["faces"], sc["colors"], sc["rem"], origin, out=out))
  miss = got["tri_id"] < 0
  assert 0.1 < miss.mean() < 0.9
  assert (got["range"][miss] == 7).all() and (got["endcolors"].reshape(-1, 3)[miss] == 7).all()   # hits only (RayTracer.cpp:72-90)
  z = _np(engine.cast(beams, sc["verts"], sc["faces"], sc["colors"], sc["rem"], origin, zero_misses=True))
  assert (z["range"][miss] == 0).all() and not z["endpoints"].reshape(-1, 3)[miss].any()
  _same(z, oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], 16, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE))


def test_host_ctrace_both_methods(engine, oracle):
  sc = synth.make_scene(11, n_side=60)
  H, W = 16, 128
  rays = oracle.create_rays(3.0, -25.0, H, W)
  rays[:W] = np.array([0, 0, 1], np.float32)
  origin = np.zeros(3, np.float32)
  ref = oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  for method in ("cast", "lbvh", "cast"):
    out = dict(endpoints=np.full(3 * H * W, 7.0, np.float32), endcolors=np.full(3 * H * W, 7, np.int32),
               range=np.full(H * W, 7.0, np.float32), endrem=np.full(H * W, 7.0, np.float32))
    got = engine.ctrace_host(rays, origin, sc["verts"].reshape(-1), sc["faces"].reshape(-1), sc["colors"].reshape(-1),
                             sc["rem"], H, outputs=out, want_ids=True, method=method)
    miss = ref["tri_id"] < 0
    assert miss[:W].all() and np.array_equal(got["tri_id"], ref["tri_id"]), method
    assert (got["range"][miss] == 7.0).all() and (got["endcolors"].reshape(-1, 3)[miss] == 7).all()
    hit = ~miss
    for k in ("range", "endrem"):
      assert np.array_equal(got[k][hit].view(np.int32), ref[k][hit].view(np.int32)), (method, k)
    for k in ("endpoints", "endcolors"):
      assert np.array_equal(got[k].reshape(-1, 3)[hit].view(np.int32), ref[k].reshape(-1, 3)[hit].view(np.int32)), (method, k)


def test_cast_full_size_equals_lbvh(engine, oracle):
  """Benchmark size (1.05 M triangles, 64 x 2048 and 128 x 2048 beams): both device paths, built on different
  structures, must return the same bits; the cast is idempotent on a reused workspace."""
  sc = synth.make_scene(1000, n_side=710)
  bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
  for sensor, origin in (("HDL-64E", (0.0, 0.0, 0.0)), ("OS1-128", (0.7, -0.4, 0.2))):
    H, W, fu, fd = synth.SENSORS[sensor]
    rays = oracle.create_rays(fu, fd, H, W)
    o = np.asarray(origin, np.float32)
    a = _np(engine.trace(bvh, rays, o, H))
    beams = engine.Beams(rays, H)
    ws = beams.workspace(sc["faces"].shape[0])
    b = _np(engine.cast(beams, sc["verts"], sc["faces"], sc["colors"], sc["rem"], o, workspace=ws))
    _same(b, a, what=sensor)
    c = _np(engine.cast(beams, sc["verts"], sc["faces"], sc["colors"], sc["rem"], o, workspace=ws))
    _same(c, b, what=sensor + " rerun")
    assert (b["tri_id"] >= 0).mean() > 0.9


def test_cast_config4_size_equals_lbvh(engine, oracle):
  """BASELINE.json configs[3] shape: a ~2 M-triangle mesh cast with the 128 x 2048 OS1-128 pattern -- both device
  paths return the same bits; every reported range is the Moller-Trumbore t of the reported triangle (host
  re-evaluation in float32 with the reference's operation order, Triangle.h:27-50)."""
  sc = synth.make_scene(4000, n_side=1000)
  assert sc["faces"].shape[0] > 1.9e6
  H, W, fu, fd = synth.SENSORS["OS1-128"]
  rays = oracle.create_rays(fu, fd, H, W)
  o = np.array([0.3, 0.1, 0.05], np.float32)
  beams = engine.Beams(rays, H)
  a = _np(engine.cast(beams, sc["verts"], sc["faces"], sc["colors"], sc["rem"], o))
  bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
  b = _np(engine.trace(bvh, rays, o, H))
  _same(a, b, what="config 4")
  hit = a["tri_id"] >= 0
  assert hit.mean() > 0.9
  f32 = np.float32
  fa = sc["faces"][a["tri_id"][hit]]
  v0, v1, v2 = (sc["verts"][fa[:, k]] for k in range(3))
  d = oracle.normalize_rays(rays, oracle.NORMALIZE_SSE).reshape(-1, 3)[hit]   # engine.DEFAULT_NORMALIZE == "sse"
  e1, e2 = (v1 - v0).astype(f32), (v2 - v0).astype(f32)
  cross = lambda p, q: np.stack([(p[:, 1] * q[:, 2]).astype(f32) - (p[:, 2] * q[:, 1]).astype(f32),
                                 (p[:, 2] * q[:, 0]).astype(f32) - (p[:, 0] * q[:, 2]).astype(f32),
                                 (p[:, 0] * q[:, 1]).astype(f32) - (p[:, 1] * q[:, 0]).astype(f32)], -1).astype(f32)
  dot = lambda p, q: (((p[:, 0] * q[:, 0]).astype(f32) + (p[:, 1] * q[:, 1]).astype(f32)).astype(f32) + (p[:, 2] * q[:, 2]).astype(f32)).astype(f32)
  inv_a = (f32(1) / dot(e1, cross(d, e2))).astype(f32)
  t = (dot(e2, cross((o[None, :] - v0).astype(f32), e1)) * inv_a).astype(f32)
  assert np.array_equal(t.view(np.int32), a["range"][hit].view(np.int32))
  assert np.array_equal(a["endcolors"].reshape(-1, 3)[hit], sc["colors"][fa[:, 0]])


def test_scan_renderer_graph_replay_equals_direct_cast(engine, oracle):
  """ScanRenderer.submit replays one captured graph per stream slot with the mesh swapped through a pinned descriptor:
  meshes of different sizes in sequence (larger, smaller, empty, larger again) on 2 slots must give what engine.cast
  gives, and what the kernel-by-kernel submission gives."""
  import torch
  from lidar_transfer_b200 import pipeline
  H, W, fu, fd = 32, 512, 10.0, -25.0
  rays = oracle.create_rays(fu, fd, H, W)
  origin = np.array([0.2, -0.1, 0.0], np.float32)
  scenes = [synth.make_scene(500 + k, n_side=n, n_boxes=4) for k, n in enumerate((90, 40, 120, 25, 90))]
  empty = dict(verts=np.zeros((3, 3), np.float32), faces=np.zeros((0, 3), np.int32), colors=np.zeros((3, 3), np.int32), rem=np.zeros(3, np.float32))
  scenes.insert(2, empty)
  max_v = max(s["verts"].shape[0] for s in scenes); max_f = max(s["faces"].shape[0] for s in scenes)
  dev = [tuple(torch.from_numpy(np.ascontiguousarray(s[k]).reshape(-1)).cuda() for k in ("verts", "faces", "colors", "rem")) for s in scenes]
  beams = engine.Beams(rays, H)
  want = [_np(engine.cast(beams, s["verts"], s["faces"], s["colors"], s["rem"], origin, zero_misses=True)) for s in scenes]
  for use_graph in (True, False):
    R = pipeline.ScanRenderer(rays, origin, H, max_v, max_f, n_streams=2, use_graph=use_graph)
    assert R.use_graph == use_graph
    for rep in range(2):
      for i, d in enumerate(dev):
        slot = R.submit(*d)
        slot.done.synchronize()
        got = {k: v.cpu().numpy() for k, v in slot.out.items()}
        _same(got, want[i], what="graph %s scene %d rep %d" % (use_graph, i, rep))
    # several in flight, then drain
    slots = [R.submit(*dev[i]) for i in (0, 3)]
    R.wait()
    for slot, i in zip(slots, (0, 3)):
      _same({k: v.cpu().numpy() for k, v in slot.out.items()}, want[i], what="in flight")
    R.close()


def test_cast_random_configurations_vs_oracle(engine, oracle):
  """A sweep nobody wrote down by hand: random beam grids (rows, columns, field of view), random origins, hostile
  soups of random size -- cast == oracle bit for bit (hypothesis drives the parameters, 20 examples)."""
  from hypothesis import given, settings, strategies as st

  @settings(max_examples=20, deadline=None)
  @given(seed=st.integers(0, 2 ** 31 - 1), H=st.integers(1, 40), W=st.integers(1, 300), n=st.integers(16, 3000),
         fu=st.floats(-10.0, 60.0), span=st.floats(1.0, 80.0), far=st.booleans())
  def run(seed, H, W, n, fu, span, far):
    rng = np.random.default_rng(seed)
    origin = (rng.normal(size=3) * (1000.0 if far else 2.0)).astype(np.float32)
    verts, faces = _hostile_mesh(rng, n, origin)
    rays = oracle.create_rays(fu, fu - span, H, W)
    _run(engine, oracle, verts, faces, rays, origin, H, lbvh=False)

  run()


def test_ctrace_beam_cache_follows_the_rays_and_threads_are_serialised(engine, oracle, vl):
  """The host entry point keeps the normalised rays + beam index of the previous call on the device: the same rays again
  must hit the cache, different rays (same count, same pointer, content edited in place) must rebuild it, and the results
  must be the oracle's either way; two Python threads calling at once are serialised by the library's mutex."""
  import ctypes
  import threading
  sc = synth.make_scene(77, n_side=70, n_boxes=6)
  H, W = 16, 128
  rays = oracle.create_rays(3.0, -25.0, H, W)
  origin = np.zeros(3, np.float32)
  args = (origin, sc["verts"].reshape(-1), sc["faces"].reshape(-1), sc["colors"].reshape(-1), sc["rem"], H)

  def stats():
    h, m = ctypes.c_longlong(0), ctypes.c_longlong(0)
    vl.vl_ctrace_cache_stats(ctypes.byref(h), ctypes.byref(m))
    return h.value, m.value
  flags = oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE
  ref = oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, flags)
  a = engine.ctrace_host(rays, *args, want_ids=True)
  h0, m0 = stats()
  b = engine.ctrace_host(rays, *args, want_ids=True)
  h1, m1 = stats()
  assert (h1, m1) == (h0 + 1, m0)                                       # same rays: index reused
  assert np.array_equal(a["tri_id"], ref["tri_id"]) and np.array_equal(b["range"].view(np.int32), ref["range"].view(np.int32))
  rays[5::3] *= np.float32(-1.0)                                         # edited in place: same buffer, new content
  ref2 = oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, flags)
  c = engine.ctrace_host(rays, *args, want_ids=True)
  h2, m2 = stats()
  assert (h2, m2) == (h1, m1 + 1)                                       # rebuilt
  assert np.array_equal(c["tri_id"], ref2["tri_id"]) and np.array_equal(c["range"].view(np.int32), ref2["range"].view(np.int32))
  assert not np.array_equal(ref["tri_id"], ref2["tri_id"])
  out = [None, None]

  def work(k):
    for _ in range(5):
      out[k] = engine.ctrace_host(rays, *args, want_ids=True)
  ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
  [t.start() for t in ts]
  [t.join() for t in ts]
  for k in range(2):
    assert np.array_equal(out[k]["tri_id"], ref2["tri_id"]) and np.array_equal(out[k]["endpoints"].view(np.int32), ref2["endpoints"].view(np.int32))
  ieee = engine.ctrace_host(rays, *args, want_ids=True, normalize="ieee")   # the portable mode: the oracle's IEEE mode
  ref3 = oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES)
  assert np.array_equal(ieee["tri_id"], ref3["tri_id"]) and np.array_equal(ieee["range"].view(np.int32), ref3["range"].view(np.int32))
  engine.ctrace_host(rays, *args, normalize="sse")                         # back to the default for the tests that follow


def test_scan_renderer_keeps_its_inputs_alive_until_the_scan_has_drained(engine, oracle):
  """ScanRenderer.submit() reads the caller's tensors on the slot's own stream, which the caching allocator does not know
  about: the slot holds references until the scan has drained, so a caller that drops its mesh right after
  submit() and immediately allocates and overwrites same-sized tensors must still get the right answer."""
  import torch
  from lidar_transfer_b200 import pipeline
  H, W = 32, 512
  rays = oracle.create_rays(10.0, -25.0, H, W)
  origin = np.zeros(3, np.float32)
  scenes = [synth.make_scene(900 + k, n_side=150, n_boxes=6) for k in range(3)]
  want = [oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE) for sc in scenes]
  mv, mf = max(s["verts"].shape[0] for s in scenes), max(s["faces"].shape[0] for s in scenes)
  rr = pipeline.ScanRenderer(rays, origin, H, mv, mf, n_streams=3)
  slots = []
  for sc in scenes:
    mesh = [torch.from_numpy(sc[k].reshape(-1)).cuda() for k in ("verts", "faces", "colors", "rem")]
    slots.append(rr.submit(*mesh))
    shapes = [(t.shape, t.dtype) for t in mesh]
    del mesh                                                             # the caller lets go at once ...
    junk = [torch.full(s, 7, dtype=d, device="cuda") for s, d in shapes]   # ... and the allocator may hand the blocks out again
    del junk
  for slot, ref in zip(slots, want):
    out = slot.result()
    assert slot.inputs is None
    assert np.array_equal(out["tri_id"].cpu().numpy(), ref["tri_id"])
    assert np.array_equal(out["range"].cpu().numpy().view(np.int32), ref["range"].view(np.int32))
  rr.close()


def _ctrace_out(n, fill=7):
  return dict(endpoints=np.full(3 * n, fill, np.float32), endcolors=np.full(3 * n, fill, np.int32),
              range=np.full(n, fill, np.float32), endrem=np.full(n, fill, np.float32))


@pytest.mark.parametrize("method", ["cast", "lbvh"])
def test_ctrace_wire_formats_do_not_change_a_bit(engine, oracle, vl, method):
  """The staging copy of ctrace packs what crosses PCIe (faces 3 x 21 bits, colours 3 bytes, end points recomputed on the
  host from the range): packed and raw calls must produce the same bytes in the caller's buffers -- hits and untouched
  misses -- and both must be the oracle's, on a ray count that is not a multiple of the height."""
  import ctypes
  sc = synth.make_scene(31, n_side=90, n_boxes=8)
  H, W = 16, 200
  rays = np.ascontiguousarray(np.concatenate([oracle.create_rays(25.0, -25.0, H, W), np.array([[0, 0, -1]] * 5, np.float32)]))
  origin = np.array([0.5, -1.25, 0.25], np.float32)
  n = rays.shape[0]
  ref = oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  args = (origin, sc["verts"].reshape(-1), sc["faces"].reshape(-1), sc["colors"].reshape(-1), sc["rem"], H)
  got, traffic = {}, {}
  try:
    for wire in (1, 0):
      vl.vl_ctrace_wire(wire)
      got[wire] = engine.ctrace_host(rays, *args, outputs=_ctrace_out(n), want_ids=True, method=method)
      a, b = ctypes.c_longlong(0), ctypes.c_longlong(0)
      vl.vl_ctrace_traffic(ctypes.byref(a), ctypes.byref(b))
      traffic[wire] = (a.value, b.value)
  finally:
    vl.vl_ctrace_wire(1)
    engine.ctrace_host(rays[:H], *args, method="cast")
  miss = ref["tri_id"] < 0
  assert 0.05 < miss.mean() < 0.95 and miss[H * W:].all()        # the five rays beyond width * height are never cast
  for wire in (1, 0):
    assert np.array_equal(got[wire]["tri_id"], ref["tri_id"])
    hit = ~miss
    for k, c in (("range", 1), ("endrem", 1), ("endpoints", 3), ("endcolors", 3)):
      x, y = got[wire][k].reshape(-1, c).view(np.int32), np.asarray(ref[k]).reshape(-1, c).view(np.int32)
      assert np.array_equal(x[hit], y[hit]), (wire, k)
      assert (got[wire][k].reshape(-1, c)[miss] == 7).all(), (wire, k)   # hits only (RayTracer.cpp:72-90)
  assert traffic[1][0] < 0.72 * traffic[0][0] and traffic[1][1] < 0.7 * traffic[0][1], traffic


def test_ctrace_arrays_that_do_not_fit_the_wire_formats_travel_raw(engine, oracle, vl):
  """Colour components outside 0 .. 255 (the reference passes them through float, RayTracer.cpp:36-48), meshes with more
  than 2^21 vertices and face indices that no 21-bit field holds must give the raw path's results."""
  from lidar_transfer_b200._lib import VlidarError, VL_EBADMESH
  sc = synth.make_scene(32, n_side=50, n_boxes=4)
  H, W = 8, 128
  rays = oracle.create_rays(3.0, -25.0, H, W)
  origin = np.zeros(3, np.float32)
  flags = oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE
  verts, faces, rem = sc["verts"], sc["faces"], sc["rem"]
  # (a) colours: negative, > 255 and beyond float's integer range
  colors = sc["colors"].copy()
  colors[::3, 0] = -5
  colors[1::3, 1] = 70000
  colors[2::3, 2] = 16777217
  ref = oracle.trace(rays, origin, verts, faces, colors, rem, H, flags)
  for method in ("cast", "lbvh"):
    got = engine.ctrace_host(rays, origin, verts.reshape(-1), faces.reshape(-1), colors.reshape(-1), rem, H, want_ids=True, method=method)
    assert np.array_equal(got["tri_id"], ref["tri_id"]) and np.array_equal(got["endcolors"], ref["endcolors"].reshape(-1)), method
    assert (got["endcolors"].reshape(-1, 3)[ref["tri_id"] >= 0].max(axis=0) > 255).any()
  # (b) more than 2^21 vertices: the faces travel as int32 triples; the mesh's own vertices sit at the END of the array
  pad = (1 << 21) + 17
  big_v = np.concatenate([np.zeros((pad, 3), np.float32), verts])
  big_c = np.concatenate([np.zeros((pad, 3), np.int32), sc["colors"]])
  big_r = np.concatenate([np.zeros(pad, np.float32), rem])
  big_f = (faces + pad).astype(np.int32)
  ref_b = oracle.trace(rays, origin, verts, faces, sc["colors"], rem, H, flags)
  got = engine.ctrace_host(rays, origin, big_v.reshape(-1), big_f.reshape(-1), big_c.reshape(-1), big_r, H, want_ids=True, method="cast")
  _same(got, ref_b, what="ctrace, 2^21 + vertices")
  # (c) a face index no 21-bit field holds (and a negative one): bad faces, skipped and reported; the others are cast
  bad_f = faces.copy()
  bad_f[3, 1] = 1 << 22
  bad_f[10, 0] = -2
  keep = np.ones(faces.shape[0], bool)
  keep[[3, 10]] = False
  ref_c = oracle.trace(rays, origin, verts, faces[keep], sc["colors"], rem, H, flags)
  ids = np.flatnonzero(keep).astype(np.int32)
  out = _ctrace_out(H * W, 0)
  tri = np.empty(H * W, np.int32)
  p = lambda a: a.ctypes.data
  rc = vl.vl_ctrace_ids(p(rays), p(origin), p(verts), p(bad_f), p(sc["colors"]), p(rem), H * W, verts.shape[0], bad_f.shape[0], H,
                        p(out["endpoints"]), p(out["endcolors"]), p(out["range"]), p(out["endrem"]), p(tri))
  assert rc == VL_EBADMESH
  want = np.where(ref_c["tri_id"] >= 0, ids[np.maximum(ref_c["tri_id"], 0)], -1)
  assert np.array_equal(tri, want) and np.array_equal(out["range"].view(np.int32), ref_c["range"].view(np.int32))
  engine.ctrace_host(rays, origin, verts.reshape(-1), faces.reshape(-1), sc["colors"].reshape(-1), rem, H)   # and the library carries on


def test_cast_takes_a_triangle_soup_without_an_index_array(engine, oracle):
  """faces=None: face f = vertices (3f, 3f+1, 3f+2), the layout the mesh extraction produces -- same bits as the explicit
  index array 0 .. 3T-1 and as the oracle, uint8 colours included; a vertex count that is not a multiple of three leaves
  the last vertices unused."""
  import torch
  sc = synth.make_scene(61, n_side=70, n_boxes=6)
  tri = sc["verts"][sc["faces"].reshape(-1)]                       # [3T, 3]
  col = sc["colors"][sc["faces"].reshape(-1)]
  rem = sc["rem"][sc["faces"].reshape(-1)]
  T = tri.shape[0] // 3
  faces = np.arange(3 * T, dtype=np.int32).reshape(T, 3)
  H, W = 32, 256
  rays = oracle.create_rays(5.0, -25.0, H, W)
  origin = np.array([0.25, 0.5, 0.1], np.float32)
  ref = oracle.trace(rays, origin, tri, faces, col, rem, H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  beams = engine.Beams(rays, H)
  a = _np(engine.cast(beams, tri, faces, col, rem, origin, check_mesh=True))
  b = _np(engine.cast(beams, tri, None, col, rem, origin, check_mesh=True))
  _same(a, ref, what="explicit faces")
  _same(b, ref, what="soup")
  assert b["n_bad_faces"] == 0 and (b["tri_id"] >= 0).mean() > 0.5
  c8 = torch.from_numpy((col & 255).astype(np.uint8)).cuda()
  c = _np(engine.cast(beams, tri, None, c8, rem, origin))
  _same(c, ref, what="soup, uint8 colours")
  pad = np.concatenate([tri, np.full((2, 3), 1e3, np.float32)])
  d = _np(engine.cast(beams, pad, None, np.concatenate([col, np.zeros((2, 3), np.int32)]), np.concatenate([rem, np.zeros(2, np.float32)]), origin))
  _same(d, ref, what="soup, two spare vertices")


@pytest.mark.parametrize("n_faces", [0, 1, 3, 1025])
@pytest.mark.parametrize("method", ["cast", "lbvh"])
def test_ctrace_tiny_and_empty_meshes(engine, oracle, vl, n_faces, method):
  """The host entry point on meshes smaller than one staging item, odd face counts (the packed face words are written two at
  a time) and no faces at all (the reference: every ray misses, nothing is written)."""
  rng = np.random.default_rng(100 + n_faces)
  sc = synth.make_scene(9, n_side=24, n_boxes=0)
  faces = np.ascontiguousarray(sc["faces"][rng.permutation(sc["faces"].shape[0])[:n_faces]])
  H, W = 8, 64
  rays = oracle.create_rays(-5.0, -40.0, H, W)
  origin = np.zeros(3, np.float32)
  out = _ctrace_out(H * W, 5)
  if n_faces == 0:
    tri = np.empty(H * W, np.int32)
    p = lambda a: a.ctypes.data
    z = np.zeros(3, np.int32)
    vl.vl_ctrace_method({"cast": 0, "lbvh": 1}[method])
    rc = vl.vl_ctrace_ids(p(rays), p(origin), p(sc["verts"]), p(z), p(sc["colors"]), p(sc["rem"]), H * W, sc["verts"].shape[0], 0, H,
                          p(out["endpoints"]), p(out["endcolors"]), p(out["range"]), p(out["endrem"]), p(tri))
    vl.vl_ctrace_method(0)
    assert rc == 0 and (tri == -1).all() and (out["range"] == 5).all() and (out["endcolors"] == 5).all()
    return
  ref = oracle.trace(rays, origin, sc["verts"], faces, sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  got = engine.ctrace_host(rays, origin, sc["verts"].reshape(-1), faces.reshape(-1), sc["colors"].reshape(-1), sc["rem"], H,
                           outputs=out, want_ids=True, method=method)
  engine.ctrace_host(rays, origin, sc["verts"].reshape(-1), faces.reshape(-1), sc["colors"].reshape(-1), sc["rem"], H, method="cast")
  hit = ref["tri_id"] >= 0
  assert np.array_equal(got["tri_id"], ref["tri_id"])
  for k, c in (("range", 1), ("endrem", 1), ("endpoints", 3), ("endcolors", 3)):
    x, y = got[k].reshape(-1, c).view(np.int32), np.asarray(ref[k]).reshape(-1, c).view(np.int32)
    assert np.array_equal(x[hit], y[hit]), k
    assert (got[k].reshape(-1, c)[~hit] == 5).all(), k
