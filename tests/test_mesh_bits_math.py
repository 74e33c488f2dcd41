"""The word arithmetic of the bit-volume mesh sweep (lidar_transfer_b200/csrc/vl_mesh.cu: k_mesh_bits, bits_at,
cube_words) restated with Python integers and checked on the CPU against the plain per-cube definition (case bit c
set when corner c -- bit 0 x, bit 1 y, bit 2 z -- is below the level; cubes need x + 1 < dx, y + 1 < dy, z + 1 < dz)
over volume shapes the GPU tests do not enumerate: z rows shorter / longer than a 32-cube word, planes that end inside
a word, single layers."""
import numpy as np
from hypothesis import given, settings, strategies as st

M32 = 0xffffffff


def _funnel_r(lo, hi, s):          # __funnelshift_r: the low 32 bits of (hi:lo) >> (s & 31)
  return (((hi << 32) | lo) >> (s & 31)) & M32


def _plane_words(bits_plane):      # k_mesh_bits: bit j of the plane -> word j >> 5, bit j & 31
  n = len(bits_plane)
  pw = (n + 31) // 32
  words = [0] * pw
  for j in np.flatnonzero(bits_plane):
    words[j >> 5] |= 1 << (int(j) & 31)
  return words, pw


def _bits_at(plane, pw, bit):      # bits_at: clamped reads
  k = bit >> 5
  return _funnel_r(plane[min(k, pw - 1)], plane[min(k + 1, pw - 1)], bit)


def _cube_words(p0, p1, pw, w, dy, dz):
  j0 = 32 * w
  c = [p0[w], p1[w], _bits_at(p0, pw, j0 + dz), _bits_at(p1, pw, j0 + dz), _bits_at(p0, pw, j0 + 1), _bits_at(p1, pw, j0 + 1),
       _bits_at(p0, pw, j0 + dz + 1), _bits_at(p1, pw, j0 + dz + 1)]
  any_, all_ = 0, M32
  for v in c:
    any_ |= v
    all_ &= v
  act = any_ & ~all_ & M32
  rows_left = (dy - 1) * dz - j0
  valid = M32 if rows_left >= 32 else (0 if rows_left <= 0 else (1 << rows_left) - 1)
  b = dz - 1 - j0 % dz
  while b < 32:
    valid &= ~(1 << b) & M32
    b += dz
  return c, act & valid


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), dx=st.integers(1, 4), dy=st.integers(1, 9), dz=st.integers(1, 70),
       density=st.sampled_from([0.02, 0.3, 0.5, 0.97]))
def test_word_sweep_equals_the_per_cube_definition(seed, dx, dy, dz, density):
  rng = np.random.default_rng(seed)
  below = rng.random((dx, dy, dz)) < density                       # tsdf < level
  yz = dy * dz
  planes = [_plane_words(below[x].reshape(-1)) for x in range(dx)]
  pw = planes[0][1]
  got = {}
  for x in range(dx - 1):
    p0, p1 = planes[x][0], planes[x + 1][0]
    for w in range(pw):
      c, act = _cube_words(p0, p1, pw, w, dy, dz)
      for b in range(32):
        if act >> b & 1:
          m = sum(((c[k] >> b) & 1) << k for k in range(8))
          got[(x, 32 * w + b)] = m
  want = {}
  for x in range(dx - 1):
    for y in range(dy - 1):
      for z in range(dz - 1):
        m = 0
        for cbit in range(8):
          if below[x + (cbit & 1), y + ((cbit >> 1) & 1), z + ((cbit >> 2) & 1)]:
            m |= 1 << cbit
        if m not in (0, 255):
          want[(x, y * dz + z)] = m
  assert got == want
  assert all(j < yz for (_, j) in got)


def _nth_active(act0, act1, excl, want):
  """nth_active: owner lane by bisection over the exclusive prefix, word by rank, bit by popcount selection."""
  o = 0
  step = 16
  while step:
    if excl[o + step] <= want:
      o += step
    step >>= 1
  rnk = want - excl[o]
  n0 = bin(act0[o]).count("1")
  h = 1 if rnk >= n0 else 0
  act = act1[o] if h else act0[o]
  rnk -= n0 if h else 0
  b, sft = 0, 16
  while sft:
    c = bin((act >> b) & ((1 << sft) - 1)).count("1")
    if rnk >= c:
      b += sft
      rnk -= c
    sft >>= 1
  return 64 * o + 32 * h + (b & 31)


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), density=st.sampled_from([0.0, 0.01, 0.05, 0.5, 1.0]))
def test_rank_selection_enumerates_the_active_cubes_of_a_unit_in_cube_order(seed, density):
  rng = np.random.default_rng(seed)
  bits = rng.random(2048) < density
  if density == 0.01:
    bits[rng.integers(0, 2048)] = True
  act0 = [int(sum(1 << b for b in range(32) if bits[64 * l + b])) for l in range(32)]
  act1 = [int(sum(1 << b for b in range(32) if bits[64 * l + 32 + b])) for l in range(32)]
  counts = [bin(a).count("1") + bin(c).count("1") for a, c in zip(act0, act1)]
  excl = [int(v) for v in np.concatenate([[0], np.cumsum(counts)[:-1]])]
  total = sum(counts)
  got = [_nth_active(act0, act1, excl, r) for r in range(total)]
  assert got == [int(j) for j in np.flatnonzero(bits)]


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n_active=st.integers(1, 3000), max_cnt=st.sampled_from([1, 2, 5]))
def test_per_triangle_emit_finds_every_triangles_cube(seed, n_active, max_cnt):
  """k_mesh_compact_bits leaves, per 256 triangle slots, the list index of the cube holding the first of them
  (cta_first[kb] for the cube whose slot range contains kb * 256); k_mesh_emit loads 256 list entries from there and
  each thread bisects for the last entry whose first slot is <= its triangle.  Every triangle must land on its cube."""
  rng = np.random.default_rng(seed)
  E = 256
  cnt = rng.integers(1, max_cnt + 1, n_active)
  first = np.concatenate([[0], np.cumsum(cnt)[:-1]])
  n_tris = int(cnt.sum())
  n_ctas = (n_tris + E - 1) // E
  cta_first = [None] * n_ctas
  for slot in range(n_active):                      # the compaction's marker rule
    tri, c = int(first[slot]), int(cnt[slot])
    kb = (tri + E - 1) // E
    if kb * E < tri + c and kb * E < n_tris:
      assert cta_first[kb] is None
      cta_first[kb] = slot
  assert all(v is not None for v in cta_first)
  owner = np.repeat(np.arange(n_active), cnt)       # the cube of every triangle
  for k in range(n_ctas):
    s_first = [int(first[cta_first[k] + i]) if cta_first[k] + i < n_active else 0xffffffff for i in range(E)]
    for t in range(min(E, n_tris - k * E)):
      T = k * E + t
      lo, step = 0, E // 2
      while step:
        if s_first[lo + step] <= T:
          lo += step
        step >>= 1
      assert cta_first[k] + lo == owner[T] and 0 <= T - s_first[lo] < cnt[owner[T]]
