"""Property tests of the oracle itself (CPU, hypothesis): the checker has to be right on inputs nobody wrote down.
  * the restated reference BVH (top-down midpoint splits, BVH.cpp:143-243) finds the same closest hit as testing every
    triangle, on random soups with degenerate and duplicated triangles;
  * the hit (range, hit mask) does not depend on the order of the triangles -- only the id of an exact tie may;
  * the projection's per-pixel winner is the nearest point whatever the order of the points (laserscan.py:373-382)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def _soup(seed, n):
  rng = np.random.default_rng(seed)
  c = rng.normal(size=(n, 1, 3)) * rng.choice([2.0, 10.0, 40.0], (n, 1, 1))
  tri = (c + rng.normal(size=(n, 3, 3)) * rng.choice([0.01, 0.3, 3.0, 20.0], (n, 1, 1))).astype(np.float32)
  k = max(1, n // 10)
  tri[:k, 2] = tri[:k, 1]                 # zero area
  if 3 * k <= n:
    tri[k:2 * k] = tri[2 * k:3 * k]       # duplicates: exact-t ties
  verts = tri.reshape(-1, 3)
  faces = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
  colors = np.zeros((3 * n, 3), np.int32); colors[:, 2] = rng.integers(1, 250, 3 * n)
  rem = rng.random(3 * n).astype(np.float32)
  rays = rng.normal(size=(64, 3)).astype(np.float32)
  origin = rng.normal(size=3).astype(np.float32) * 0.5
  return verts, faces, colors, rem, rays, origin


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 300))
def test_bvh_restatement_equals_brute_force(oracle, seed, n):
  verts, faces, colors, rem, rays, origin = _soup(seed, n)
  a = oracle.trace(rays, origin, verts, faces, colors, rem, 8, oracle.MIN_ID_TIES)
  b = oracle.trace(rays, origin, verts, faces, colors, rem, 8, oracle.BRUTE_FORCE)
  for k in ("tri_id", "range", "endpoints", "endcolors", "endrem"):
    assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), k


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(2, 200))
def test_hit_is_independent_of_triangle_order(oracle, seed, n):
  verts, faces, colors, rem, rays, origin = _soup(seed, n)
  perm = np.random.default_rng(seed + 1).permutation(n)
  a = oracle.trace(rays, origin, verts, faces, colors, rem, 8, oracle.MIN_ID_TIES)
  b = oracle.trace(rays, origin, verts, faces[perm], colors, rem, 8, oracle.MIN_ID_TIES)
  assert np.array_equal(a["tri_id"] >= 0, b["tri_id"] >= 0)
  assert np.array_equal(a["range"].view(np.int32), b["range"].view(np.int32))
  assert np.array_equal(a["endpoints"].view(np.int32), b["endpoints"].view(np.int32))
  hit = a["tri_id"] >= 0
  back = perm[b["tri_id"][hit]]           # the permuted run's winner, in original numbering
  same = back == a["tri_id"][hit]
  # where the winners differ they are an exact tie: the same range from a different triangle
  if not same.all():
    assert np.array_equal(a["range"][hit][~same].view(np.int32), b["range"][hit][~same].view(np.int32))


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 4000))
def test_projection_keeps_the_nearest_point_whatever_the_order(oracle, seed, n):
  rng = np.random.default_rng(seed)
  H, W, fu, fd = 8, 32, 10.0, -30.0
  pts = rng.normal(size=(n, 3)) * rng.choice([1.0, 10.0, 50.0], (n, 1))
  pts[rng.random(n) < 0.05] = 0.0                                  # depth 0: dropped (laserscan.py:307-309)
  pts[: n // 4] = pts[n // 4: 2 * (n // 4)] * (1.0 + 1e-12)        # near-ties in float64 that are ties in float32
  rem = rng.random(n).astype(np.float32)
  lab = rng.integers(0, 300, n).astype(np.uint32)
  a = oracle.project(pts, rem, lab, fu, fd, H, W)
  chk = oracle.project_numpy(pts, rem, lab, fu, fd, H, W)
  for k in ("range_image", "index", "proj_label", "proj_remissions"):
    assert np.array_equal(a[k], chk[k]), k                          # C restatement == numpy restatement
  perm = rng.permutation(n)
  b = oracle.project(pts[perm], rem[perm], lab[perm], fu, fd, H, W)
  assert np.array_equal(a["range_image"].view(np.int32), b["range_image"].view(np.int32))   # the minimum is order-free
  assert np.array_equal(a["index"] >= 0, b["index"] >= 0)
  # the winning depth is the smallest float32-rounded depth of the pixel's points (ties: the loop's order decides who)
  kept = np.flatnonzero(a["keep"]) if "keep" in a else None
  if kept is not None and a["n_kept"] > 0:
    d = np.linalg.norm(pts[kept], 2, axis=1).astype(np.float32)
    filled = a["index"] >= 0
    assert np.array_equal(d[a["index"][filled]].view(np.int32), a["range_image"][filled].view(np.int32))


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 400), n_ba=st.integers(1, 40), remove=st.booleans())
def test_beam_angle_snapping_c_restatement_equals_the_python_loop(oracle, seed, n, n_ba, remove):
  """vlo_project_snap vs the line-for-line numpy restatement of laserscan.py:321-327 (both pinned to the reference's own
  Python by golden_beams_v1.npz) on random points and random lists with duplicates and exact mid-points (argmin ties:
  the first entry wins)."""
  rng = np.random.default_rng(seed)
  pts = rng.normal(0, 10, (n, 3))
  pts[rng.random(n) < 0.05] = 0.0
  ba = np.sort(rng.uniform(-0.6, 0.2, n_ba))
  ba = np.concatenate([ba, ba[:1], ba[-1:]]).tolist()      # duplicates
  if n_ba >= 2:                                           # a point whose pitch is next to the half-way mark of two entries
    # (1e-7 rad off it: numpy's SIMD arcsin and libm's asin differ by an ulp on AVX-512 hosts, which would decide an
    # exact mid-point differently in the two restatements; exact argmin ties are exercised by the duplicates above)
    mid = 0.5 * (ba[0] + ba[1]) + 1e-7
    pts[0] = [np.cos(mid) * 7.0, 0.0, np.sin(mid) * 7.0]
  rem = rng.random(n).astype(np.float32)
  lab = rng.integers(0, 260, n).astype(np.uint32)
  a = oracle.project(pts, rem, lab, 3.0, -25.0, 16, 64, remove=remove, beam_angles=ba)
  b = oracle.project_numpy(pts, rem, lab, 3.0, -25.0, 16, 64, remove=remove, beam_angles=ba)
  assert a["n_kept"] == b["n_kept"] and np.array_equal(a["keep"], b["keep"])
  for k in ("index", "proj_label"):
    assert np.array_equal(a[k], b[k]), k
  for k in ("range_image", "proj_remissions"):
    assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), k


@settings(max_examples=200, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 40), spread=st.sampled_from([0.0, 1e-9, 1e-7, 1e-3]))
def test_one_atomic_min_key_reproduces_the_sequential_pixel_loop(seed, n, spread):
  """vl_project.cu replaces the reference's sequential per-point loop (laserscan.py:373-382: a point wins a pixel when
  its float64 depth is below the float32 value stored there, or the pixel is empty) by ONE 64-bit atomicMin per point on
  key = float32(depth) bits << 32 | (depth < float32(depth) ? 0x7fffffff - i : 0x80000000 | i).  The winner of the
  minimum key must be the loop's winner for every arrival order -- including float64 depths that collapse onto one
  float32 value, where the loop's outcome depends on which side of that value each depth lies."""
  rng = np.random.default_rng(seed)
  base = rng.uniform(0.5, 80.0)
  depth = base * (1.0 + spread * rng.integers(-3, 4, n))          # float64, many exact ties
  # the reference's loop for one pixel
  stored, winner = np.float32(0.0), -1
  for i in range(n):
    if depth[i] < stored or winner == -1:
      stored, winner = np.float32(depth[i]), i
  # the key
  keys = []
  for i in range(n):
    R = np.float32(depth[i])
    lo = (0x7fffffff - i) if depth[i] < np.float64(R) else (0x80000000 | i)
    keys.append((int(R.view(np.uint32)) << 32) | lo)
  k = min(keys)
  got = (0x7fffffff - (k & 0xffffffff)) if (k & 0xffffffff) < 0x80000000 else (k & 0x7fffffff)
  assert got == winner and np.uint32(k >> 32) == stored.view(np.uint32)
