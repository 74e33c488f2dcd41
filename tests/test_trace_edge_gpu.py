"""Edge cases of rows (i)+(ii) the reference's contract implies and size-independent
properties at the benchmark's full size.  CUDA through the C ABI vs the oracle, bit for bit."""
import numpy as np
import pytest

from lidar_transfer_b200 import synth

pytestmark = pytest.mark.gpu
KEYS = ("tri_id", "range", "endpoints", "endcolors", "endrem")


def _np(out):
  return {k: v.cpu().numpy() for k, v in out.items()}


def _same(got, ref, keys=KEYS):
  for k in keys:
    assert np.array_equal(np.asarray(got[k]).view(np.int32), np.asarray(ref[k]).view(np.int32)), k


def _run(engine, oracle, verts, faces, rays, origin, H, colors=None, rem=None):
  verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
  faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
  if colors is None:
    colors = np.zeros(verts.shape, np.int32); colors[:, 2] = 40 + np.arange(verts.shape[0]) % 7
  if rem is None:
    rem = (np.arange(verts.shape[0]) % 13 / 13.0).astype(np.float32)
  rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 3)
  origin = np.asarray(origin, np.float32)
  ref = oracle.trace(rays, origin, verts, faces, colors, rem, H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  bvh = engine.Bvh(verts, faces, colors, rem)
  got = _np(engine.trace(bvh, rays, origin, H))
  _same(got, ref)
  bf = _np(engine.trace_bruteforce(verts, faces, colors, rem, rays, origin, H))
  _same(bf, ref)
  return got, ref


def test_shared_edges_vertices_and_parallel_rays(engine, oracle):
  """Rays aimed exactly at shared edges / vertices (ties -> smaller face index), rays parallel to a triangle's
  plane (|a| < eps -> miss, Triangle.h:33-35) and rays starting on the surface (t < eps -> miss)."""
  g = np.arange(5, dtype=np.float32)
  X, Y = np.meshgrid(g, g, indexing="ij")
  verts = np.stack([X, Y, np.full_like(X, -2.0)], -1).reshape(-1, 3)
  idx = np.arange(25).reshape(5, 5)
  a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
  faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
  origin = np.array([2.0, 2.0, 0.0], np.float32)
  targets = [(x, y, -2.0) for x in (0.0, 0.5, 1.0, 2.0, 2.5, 4.0) for y in (0.0, 1.0, 1.5, 2.0, 3.0, 4.0)]  # vertices, edges, interiors
  rays = [np.array(t) - origin for t in targets]
  rays += [(1.0, 0.0, 0.0), (0.0, -1.0, 0.0), (1.0, 1.0, 0.0)]  # parallel to the plane: miss
  rays += [(0.0, 0.0, 1.0), (0.0, 0.0, -1.0)]                   # straight up (miss) and straight down (axis-aligned: NaN lanes in the slab test)
  rays += [(0.0, 0.0, 0.0)]                                     # zero direction: normalises to NaN, must miss
  while len(rays) % 4:
    rays.append((0.3, -0.2, -1.0))
  got, ref = _run(engine, oracle, verts, faces, np.array(rays, np.float32), origin, 4)
  # (rays aimed at the outer border of the grid may round to u + v > 1: whatever the reference arithmetic says, holds)
  assert (got["tri_id"][:36] >= 0).sum() >= 30 and (got["tri_id"][36:39] == -1).all() and got["tri_id"][39] == -1 and got["tri_id"][40] >= 0
  assert got["tri_id"][41] == -1
  # a ray that starts ON the surface does not hit it (t < 1e-6)
  got2, _ = _run(engine, oracle, verts, faces, np.array([[0.1, 0.2, -1.0]] * 4, np.float32), np.array([1.3, 1.6, -2.0], np.float32), 2)
  assert (got2["tri_id"] == -1).all()


def test_degenerate_triangles_and_duplicates(engine, oracle):
  """Zero-area triangles (never hit), duplicated triangles (exact-t tie -> smaller face index), coincident
  centroids (identical Morton keys: the index tie-break of the radix tree)."""
  sc = synth.make_scene(5, n_side=12, n_boxes=2)
  verts, faces = sc["verts"], sc["faces"]
  dup = faces[::3].copy()
  zero = np.stack([faces[::5, 0], faces[::5, 0], faces[::5, 1]], -1)  # two equal corners
  line = np.stack([faces[::7, 0], faces[::7, 1], faces[::7, 1]], -1)
  faces2 = np.concatenate([zero, faces, dup, line, dup])
  rays = oracle.create_rays(3.0, -25.0, 8, 64)
  got, ref = _run(engine, oracle, verts, faces2, rays, np.zeros(3, np.float32), 8, sc["colors"], sc["rem"])
  hit = got["tri_id"] >= 0
  assert hit.mean() > 0.5
  n0 = zero.shape[0]
  assert ((got["tri_id"][hit] >= n0) & (got["tri_id"][hit] < n0 + faces.shape[0])).all()  # never a degenerate, never a later duplicate


@pytest.mark.parametrize("n_faces", [0, 1, 2, 3, 4, 5, 8, 9, 511, 512, 513, 1025])
def test_tiny_meshes_and_cta_boundaries(engine, oracle, n_faces):
  """0..5 triangles (single leaf / single node), and meshes around the 512-triangle CTA of the hierarchy kernel."""
  rng = np.random.default_rng(n_faces)
  sc = synth.make_scene(9, n_side=24, n_boxes=0)
  pick = rng.permutation(sc["faces"].shape[0])[:n_faces]
  faces = sc["faces"][pick]
  rays = oracle.create_rays(-5.0, -40.0, 8, 64)
  got, ref = _run(engine, oracle, sc["verts"], faces, rays, np.zeros(3, np.float32), 8, sc["colors"], sc["rem"])
  if n_faces == 0:
    assert (got["tri_id"] == -1).all() and (got["range"] == 0).all()


def test_ragged_ray_count_and_all_miss(engine, oracle):
  """n_rays % height != 0: width = n_rays // height, the trailing rays are never cast (RayTracer.cpp:56); a ray
  set that misses everything leaves every output zero."""
  sc = synth.make_scene(21, n_side=20, n_boxes=3)
  rays = oracle.create_rays(3.0, -25.0, 7, 33)[:7 * 33 - 5]  # 226 rays, height 7 -> width 32, 2 rays left over
  got, ref = _run(engine, oracle, sc["verts"], sc["faces"], rays, np.zeros(3, np.float32), 7, sc["colors"], sc["rem"])
  assert (got["tri_id"][7 * 32:] == -1).all() and (got["range"][7 * 32:] == 0).all()
  up = np.tile(np.array([[0.1, 0.05, 1.0]], np.float32), (64, 1))
  got, ref = _run(engine, oracle, sc["verts"], sc["faces"], up, np.array([0, 0, 60.0], np.float32), 8, sc["colors"], sc["rem"])
  assert (got["tri_id"] == -1).all() and not got["endpoints"].any() and not got["endcolors"].any()


def test_origin_inside_scene_and_large_coordinates(engine, oracle):
  """Origin inside the mesh's bounding box away from 0 (box-entry distances negative), and a scene translated
  to UTM-like coordinates (1e5 m: float32 spacing 8 mm) -- the conservative box padding must still hold."""
  sc = synth.make_scene(31, n_side=40, n_boxes=6)
  rays = oracle.create_rays(10.0, -30.0, 16, 64)
  _run(engine, oracle, sc["verts"], sc["faces"], rays, np.array([12.5, -7.25, 0.4], np.float32), 16, sc["colors"], sc["rem"])
  shift = np.array([1.0e5, -2.0e5, 300.0], np.float32)
  got, _ = _run(engine, oracle, sc["verts"] + shift, sc["faces"], rays, shift + np.array([0.5, 0.25, 0.1], np.float32), 16,
                sc["colors"], sc["rem"])
  assert (got["tri_id"] >= 0).mean() > 0.5


def test_bad_face_index_is_reported_not_undefined(engine):
  from lidar_transfer_b200._lib import VlidarError, VL_EBADMESH
  sc = synth.make_scene(2, n_side=10, n_boxes=0)
  faces = sc["faces"].copy()
  faces[7, 1] = sc["verts"].shape[0] + 5
  faces[11, 0] = -1
  bvh = engine.Bvh(sc["verts"], faces, sc["colors"], sc["rem"])
  with pytest.raises(VlidarError) as e:
    bvh.status()
  assert e.value.code == VL_EBADMESH and "2 face" in str(e.value)
  out = _np(engine.trace(bvh, np.array([[0, 0, -1.0]] * 4, np.float32), np.zeros(3, np.float32), 2))  # still traceable
  assert not np.isin(out["tri_id"], [7, 11]).any()


def test_full_size_properties(engine, oracle):
  """At the benchmark's size (1.05 M triangles, 64 x 2048 rays), where the oracle takes too long to run per test:
  (a) the build is deterministic and the trace idempotent, (b) the BVH result equals the brute-force kernel on a
  ray subset, (c) every reported hit is consistent -- re-evaluating Moller-Trumbore for the reported triangle on
  the host gives the reported range bit for bit, and endpoints = o + d t."""
  import torch
  sc = synth.make_scene(1000, n_side=710)
  H, W = 64, 2048
  rays = oracle.create_rays(3.0, -25.0, H, W)
  origin = np.zeros(3, np.float32)
  bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
  a = _np(engine.trace(bvh, rays, origin, H))
  bvh2 = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
  b = _np(engine.trace(bvh2, rays, origin, H))
  _same(a, b)
  # header: n_tris, root reference, bad faces, climb depth, scene bounds (the header's padding and the node slots that
  # no sub-tree of > 4 triangles owns are never written, so they hold whatever the allocation held)
  # (bytes 12..16 = the deepest climb, a statistic that depends on which child reaches a node second: not compared)
  assert torch.equal(bvh.blob[:12], bvh2.blob[:12]) and torch.equal(bvh.blob[16:40], bvh2.blob[16:40])
  assert (a["tri_id"] >= 0).mean() > 0.99
  sub = np.arange(0, H * W, 997)[: 8 * 16]
  bf = _np(engine.trace_bruteforce(sc["verts"], sc["faces"], sc["colors"], sc["rem"], rays[sub], origin, 8))
  for k in KEYS:
    full = a[k].reshape(H * W, -1)[sub].reshape(bf[k].shape)
    assert np.array_equal(full.view(np.int32), bf[k].view(np.int32)), k
  # host re-evaluation of the reported triangle (oracle on a one-triangle mesh per ray would be slow: vectorised numpy
  # with explicit float32 roundings in the reference's operation order, Triangle.h:27-50)
  f32 = np.float32
  tid = a["tri_id"]
  hit = tid >= 0
  fa = sc["faces"][tid[hit]]
  v0, v1, v2 = (sc["verts"][fa[:, k]] for k in range(3))
  d = oracle.normalize_rays(rays, oracle.NORMALIZE_SSE).reshape(-1, 3)[hit]   # engine.DEFAULT_NORMALIZE == "sse"
  e1, e2 = (v1 - v0).astype(f32), (v2 - v0).astype(f32)
  cross = lambda p, q: np.stack([(p[:, 1] * q[:, 2]).astype(f32) - (p[:, 2] * q[:, 1]).astype(f32),
                                 (p[:, 2] * q[:, 0]).astype(f32) - (p[:, 0] * q[:, 2]).astype(f32),
                                 (p[:, 0] * q[:, 1]).astype(f32) - (p[:, 1] * q[:, 0]).astype(f32)], -1).astype(f32)
  dot = lambda p, q: (((p[:, 0] * q[:, 0]).astype(f32) + (p[:, 1] * q[:, 1]).astype(f32)).astype(f32) + (p[:, 2] * q[:, 2]).astype(f32)).astype(f32)
  h = cross(d, e2)
  inv_a = (f32(1) / dot(e1, h)).astype(f32)
  s = (origin[None, :] - v0).astype(f32)
  q = cross(s, e1)
  t = (dot(e2, q) * inv_a).astype(f32)
  assert np.array_equal(t.view(np.int32), a["range"][hit].view(np.int32))
  ep = (origin[None, :] + (d * t[:, None]).astype(f32)).astype(f32)
  assert np.array_equal(ep.view(np.int32), a["endpoints"].reshape(-1, 3)[hit].view(np.int32))
  assert np.array_equal(a["endcolors"].reshape(-1, 3)[hit], sc["colors"][fa[:, 0]])
