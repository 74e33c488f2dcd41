"""The geometric claims behind the scene-streaming cast's conservative (azimuth, sin elevation) rectangle
(lidar_transfer_b200/csrc/vl_cast.cu: tri_setup, pseudo_yaw, wrap_2), restated in numpy float32 and checked on the CPU:
every direction from the sensor origin to a point of the triangle -- interior, edges, vertices -- lies inside the
rectangle.  Claims: azimuth is monotonic along an edge (the interval is spanned by the vertex azimuths unless they do
not fit in a half circle = the z axis pierces the triangle), sin(elevation) exceeds its vertex values by at most
0.31 chord^2 along an edge and has no interior extremum except at a pole, the pads cover the rounding of v - o.
The GPU tests check the consequence (bit-identical hits vs brute force) on the device; this covers the bound itself,
triangle by triangle, over scales and positions nobody enumerated."""
import numpy as np
from hypothesis import given, settings, strategies as st

F = np.float32
K_PAD0 = F(2e-5)


def _pseudo_yaw(y, x):
  m = np.abs(x) + np.abs(y)
  with np.errstate(invalid="ignore", divide="ignore"):
    t = np.where(m > 0, y / m, F(0)).astype(F)
  return np.where(x >= 0, t, np.where(y >= 0, F(2) - t, F(-2) - t)).astype(F)


def _wrap_2(x):
  return (x - F(4) * np.rint(x * F(0.25))).astype(F)


def _rectangle(a, b, c, o):
  """tri_setup's interval part for n triangles: slo, shi, all_yaw, ymid, yhalf (float32, the kernel's operation order)."""
  p0, p1, p2 = (a - o).astype(F), (b - o).astype(F), (c - o).astype(F)
  q = [np.sum(p * p, axis=1, dtype=F) for p in (p0, p1, p2)]
  qmin, qmax = np.minimum(q[0], np.minimum(q[1], q[2])), np.maximum(q[0], np.maximum(q[1], q[2]))
  odd = ~((qmin > 1e-30) & (qmax < 1e30))
  with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
    r = [(F(1) / np.sqrt(qq)).astype(F) for qq in q]
    amax = np.max(np.abs(np.concatenate([a, b, c, np.broadcast_to(o, a.shape)], axis=1)), axis=1).astype(F)
    E = amax * F(2.4e-7)
    u = [(p * rr[:, None]).astype(F) for p, rr in zip((p0, p1, p2), r)]
    ch = lambda s, t: np.sum((s - t) * (s - t), axis=1, dtype=F)
    bulge = F(0.31) * np.maximum(ch(u[0], u[1]), np.maximum(ch(u[1], u[2]), ch(u[2], u[0])))
    rmax = np.maximum(r[0], np.maximum(r[1], r[2]))
    pad_s = K_PAD0 + F(2) * E * rmax
    uz = np.stack([u[0][:, 2], u[1][:, 2], u[2][:, 2]])
    slo, shi = uz.min(0) - bulge - pad_s, uz.max(0) + bulge + pad_s
    e2 = F(2) * E
    px, py, pz = np.stack([p0[:, 0], p1[:, 0], p2[:, 0]]), np.stack([p0[:, 1], p1[:, 1], p2[:, 1]]), np.stack([p0[:, 2], p1[:, 2], p2[:, 2]])
    near_axis = (px.min(0) <= e2) & (px.max(0) >= -e2) & (py.min(0) <= e2) & (py.max(0) >= -e2)
    h = px * px + py * py
    hmin = h.min(0)
    all_yaw = ~(hmin > 1e-30)
    pad_y = K_PAD0 + F(2) * E * (F(1) / np.sqrt(np.maximum(hmin, F(1e-30)))).astype(F)
    y0 = _pseudo_yaw(p0[:, 1], p0[:, 0])
    d1, d2 = _wrap_2(_pseudo_yaw(p1[:, 1], p1[:, 0]) - y0), _wrap_2(_pseudo_yaw(p2[:, 1], p2[:, 0]) - y0)
    lo_d, hi_d = np.minimum(F(0), np.minimum(d1, d2)), np.maximum(F(0), np.maximum(d1, d2))
    all_yaw = all_yaw | ~((hi_d - lo_d) + F(2) * pad_y < F(2) - F(1e-3))
    pole = all_yaw & near_axis
    shi = np.where(pole & (pz.max(0) > 0), F(2), shi)
    slo = np.where(pole & (pz.min(0) < 0), F(-2), slo)
  slo, shi, all_yaw = np.where(odd, F(-2), slo), np.where(odd, F(2), shi), all_yaw | odd
  return slo, shi, all_yaw, (y0 + F(0.5) * (lo_d + hi_d)).astype(F), (F(0.5) * (hi_d - lo_d) + pad_y).astype(F)


def _triangles(rng, n, o):
  """A mix of scales and positions around the origin o, plus the awkward families."""
  centre = rng.normal(size=(n, 3)) * rng.choice([0.3, 2.0, 10.0, 50.0], (n, 1))
  size = rng.choice([0.005, 0.05, 0.5, 5.0, 40.0], (n, 1, 1))
  tri = centre[:, None, :] + rng.normal(size=(n, 3, 3)) * size
  k = n // 6
  tri[:k, :, :2] = rng.normal(size=(k, 3, 2)) * rng.choice([0.01, 1.0, 20.0], (k, 1, 1))      # around the z axis
  tri[:k, :, 2] = rng.normal(size=(k, 3)) * 3 + rng.choice([-5, 0.0, 5], (k, 1))
  tri[k:2 * k, :, 0] = -np.abs(tri[k:2 * k, :, 0]) - 0.1                                       # across the yaw wrap (x < 0, y ~ 0)
  tri[k:2 * k, :, 1] = rng.normal(size=(k, 3)) * rng.choice([0.01, 1.0], (k, 1))
  tri[2 * k:3 * k, 2] = tri[2 * k:3 * k, 1] + rng.normal(size=(k, 3)) * 1e-4                   # slivers
  tri[3 * k:4 * k, 0] = 0.0                                                                    # a vertex AT the sensor origin
  return (tri + o).astype(F)


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), offset=st.sampled_from([0.0, 0.5, 100.0]))
def test_every_direction_to_a_point_of_the_triangle_lies_in_its_rectangle(seed, offset):
  rng = np.random.default_rng(seed)
  o = (rng.normal(size=3) * offset).astype(F)
  n, m = 1500, 24
  tri = _triangles(rng, n, o.astype(np.float64))
  a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
  slo, shi, all_yaw, ymid, yhalf = _rectangle(a, b, c, o)
  # points of the triangle: interior, the three edges, the three vertices
  w = rng.dirichlet([1, 1, 1], (n, m))
  w[:, :6] = np.where(np.eye(3)[rng.integers(0, 3, (n, 6))] > 0, 0.0, w[:, :6]); w[:, :6] /= w[:, :6].sum(-1, keepdims=True)
  w[:, 6:9] = np.eye(3)[None]
  pts = np.einsum("nmk,nkd->nmd", w, tri.astype(np.float64))
  d = pts - o.astype(np.float64)
  norm = np.linalg.norm(d, axis=-1, keepdims=True)
  ok = norm[..., 0] > 1e-9                              # a point AT the origin has no direction
  with np.errstate(invalid="ignore", divide="ignore"):
    d = (d / norm).astype(F)
  s = d[..., 2]
  yaw = _pseudo_yaw(d[..., 1], d[..., 0])
  in_s = (s >= slo[:, None]) & (s <= shi[:, None])
  in_y = all_yaw[:, None] | (np.abs(_wrap_2(yaw - ymid[:, None])) <= yhalf[:, None])
  # a direction along the z axis has no azimuth (the kernel gives such a beam yaw 0 and the triangle test decides):
  # it can only occur for a triangle the z axis touches, whose rectangle has every azimuth
  axis_dir = (np.abs(d[..., 0]) + np.abs(d[..., 1])) < 1e-6
  bad = ok & ~(in_s & (in_y | axis_dir))
  assert not bad.any(), (int(bad.sum()), np.argwhere(bad)[:3].tolist())
  assert (axis_dir & ok & ~all_yaw[:, None]).sum() == 0
  assert all_yaw.mean() < 0.6                            # and the rectangle is not vacuous: most triangles get a real interval
