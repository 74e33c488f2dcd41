"""The arithmetic behind the SPARSE TSDF volumes (lidar_transfer_b200/csrc/vl_tsdf.cu: k_tsdf_rows, k_tsdf_range_table,
k_tsdf_hull), restated in numpy float32 and checked on the CPU: the hull [z_lo, z_hi] a z column gets from a range image
must contain EVERY voxel of that column that the reference's kernel string (fusion_lidar.py:119-227) would change in a
never-written volume -- whatever the last bits of its norm3df / asinf are -- because voxels outside the hull are never
looked at again.  It must also be worth having (a small part of the volume).  The GPU tests hold the kernels to the
reference's own CUDA kernel on ten configurations; this test covers the space between them, and fails when a margin is
mutated (rows narrower than they are, a shell half as deep)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from test_tsdf_shell_math import F, K_ASIN_ERR, K_DEPTH_REL, _fma, _scene

RANGE_BINS = 256


def _reference_changes(rng, x, y, z, px, depth_im, color_im, H, fov_up, fov_down, trunc):
  """The kernel string's decision on a fresh volume, with depth and pitch pushed per voxel by up to 0.9 of the error the
  sweep tolerates (as in test_tsdf_shell_math: the claim has to hold for ANY implementation of norm3df / asinf)."""
  fov_tot = F(abs(fov_up) + abs(fov_down))
  with np.errstate(invalid="ignore", divide="ignore"):
    d64 = np.sqrt(x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2 + z.astype(np.float64) ** 2)
    depth = (d64 * (1 + 0.45 * K_DEPTH_REL * rng.uniform(-1, 1, d64.shape))).astype(F)
    pitch = (np.arcsin(np.clip(z.astype(np.float64) / d64, -1, 1)) + 0.9 * K_ASIN_ERR * rng.uniform(-1, 1, d64.shape)).astype(F)
    in_fov = ~((pitch > fov_up) | (pitch < fov_down))
    proj_y = (1.0 - (pitch.astype(np.float64) + abs(np.float64(fov_down))) / np.float64(fov_tot)).astype(F) * F(H)
    row = np.clip(np.nan_to_num(np.floor(proj_y), nan=0.0), 0, H - 1).astype(int)
    dv = depth_im[row, px]
    diff = dv - depth
    return in_fov & (dv != 0) & ~(diff < -trunc) & ((color_im[row, px] == 0) | (np.minimum(F(1), diff / trunc) < 0))


def _hulls(x2d, y2d, px2d, depth_im, color_im, H, W, fov_up, fov_down, trunc, oz, vox, dz, xy_max, mutate=None):
  """k_tsdf_shell + k_tsdf_rows + k_tsdf_range_table + k_tsdf_hull for every column (x2d, y2d: column positions)."""
  fov_tot = float(F(abs(fov_up) + abs(fov_down)))
  fov_rad = abs(float(fov_up)) + abs(float(fov_down))
  fd_abs = abs(float(fov_down))
  eps_row = 1.02 * K_ASIN_ERR * H / fov_rad + 2e-4 + 4e-7 * H
  e_p = (eps_row + 1e-2) * fov_rad / H + 2e-5
  if mutate == "no_pitch_slack":
    e_p = -0.45 * fov_rad / H      # rows narrower than they are
  lo = np.where(depth_im == 0, np.inf, np.where(color_im == 0, -np.inf, depth_im)).astype(F)
  hi = np.where(depth_im == 0, -np.inf, depth_im + (trunc * F(0.5) if mutate == "half_shell" else trunc)).astype(F)
  junk = (depth_im != 0) & ~np.isfinite(depth_im)
  lo[junk], hi[junk] = -np.inf, np.inf
  r = np.arange(H)
  p_hi = fov_tot * (1.0 - r / H) - fd_abs + e_p
  p_lo = fov_tot * (1.0 - (r + 1) / H) - fd_abs - e_p
  t_lo, t_hi = np.tan(p_lo), np.tan(p_hi)
  T_lo = (t_lo - np.abs(t_lo) * 1e-6 - 1e-7).astype(F)
  T_hi = (t_hi + np.abs(t_hi) * 1e-6 + 1e-7).astype(F)
  a_max = np.maximum(np.abs(p_lo), np.abs(p_hi))
  a_min = np.where((p_lo <= 0) & (p_hi >= 0), 0.0, np.minimum(np.abs(p_lo), np.abs(p_hi)))
  c_min, c_max = (np.cos(a_max) * (1 - 1e-6)).astype(F), (np.cos(a_min) * (1 + 1e-6)).astype(F)
  # range table: per (image column, bin of horizontal distance) the rows whose shell reaches it
  bin_inv = F((RANGE_BINS - 1) / (xy_max * 1.001 + 1e-3))
  rmin = np.full((W, RANGE_BINS), 1 << 30)
  rmax1 = np.zeros((W, RANGE_BINS), int)
  slack = 0 if mutate == "no_bin_slack" else 1
  with np.errstate(invalid="ignore", over="ignore"):
    for rr in range(H):
      for c in range(W):
        if hi[rr, c] < lo[rr, c]:
          continue
        l, h = max(float(lo[rr, c]), 0.0) * (1 - 3e-5), float(hi[rr, c]) * (1 + 3e-5)
        if not (h >= 0):
          continue
        b0 = max(0, int(np.floor((l * c_min[rr] - 1e-3) * bin_inv)) - slack)
        top = (h * c_max[rr] + 1e-3) * bin_inv
        b1 = int(np.floor(top)) + slack if top < RANGE_BINS - 1 else RANGE_BINS - 1
        b1 = min(b1, RANGE_BINS - 1)
        rmin[c, b0:b1 + 1] = np.minimum(rmin[c, b0:b1 + 1], rr)
        rmax1[c, b0:b1 + 1] = np.maximum(rmax1[c, b0:b1 + 1], rr + 1)
  xy2 = _fma(x2d, x2d, (y2d * y2d).astype(F))
  rho = np.sqrt(xy2).astype(F)
  z_bot, z_top = F(oz - vox), F(oz + dz * vox)
  m = rho * F(2e-6) + F(1e-5)
  zmin = np.full(rho.shape, np.inf, F)
  zmax = np.full(rho.shape, -np.inf, F)
  b = np.minimum(RANGE_BINS - 1, (rho * bin_inv).astype(int))
  first, last = rmin[px2d, b], rmax1[px2d, b] - 1
  with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
    for rr in range(H):
      use = (first <= rr) & (rr <= last)
      shl, shh = lo[rr, px2d], hi[rr, px2d]
      use &= ~(shh < shl)
      za, zb = rho * T_lo[rr] - m, rho * T_hi[rr] + m
      use &= ~(zb < z_bot) & ~(za > z_top)
      za, zb = np.maximum(za, z_bot), np.minimum(zb, z_top)
      h2 = shh * F(1 + 3e-5)
      B = h2 * h2 - xy2
      use &= ~(B < 0)
      l2 = shl * F(1 - 3e-5)
      A = np.where(l2 > 0, l2 * l2 - xy2, F(-1))
      sB = np.sqrt(np.maximum(B, 0)) * F(1 + 1e-6) + F(1e-6)
      a, bb = np.maximum(za, -sB), np.minimum(zb, sB)
      use &= ~(a > bb)
      sA = np.sqrt(np.maximum(A, 0)) * F(1 - 1e-6) - F(1e-6)
      hole = (A > 0) & (sA > 0)
      bl, ar = np.minimum(bb, -sA), np.maximum(a, sA)
      left, right = a <= bl, ar <= bb
      use &= ~(hole & ~left & ~right)
      a2 = np.where(hole, np.where(left, a, ar), a)
      b2 = np.where(hole, np.where(right, bb, bl), bb)
      zmin = np.where(use, np.minimum(zmin, a2), zmin)
      zmax = np.where(use, np.maximum(zmax, b2), zmax)
  pad = 0 if mutate == "no_voxel_slack" else 1
  with np.errstate(invalid="ignore", over="ignore"):
    zlo = np.maximum(0, np.floor(np.where(np.isfinite(zmin), (zmin - F(oz)) / F(vox), 0)).astype(int) - pad)
    zhi = np.minimum(dz - 1, np.ceil(np.where(np.isfinite(zmax), (zmax - F(oz)) / F(vox), -1)).astype(int) + pad)
  empty = ~(zmin <= zmax) | (zlo > zhi)
  zlo[empty], zhi[empty] = 1, 0
  near_axis = ~(rho > 1e-3)
  zlo[near_axis], zhi[near_axis] = 0, dz - 1
  return zlo, zhi


def _case(seed, H, fov, vox, zero_frac, mutate=None):
  rng = np.random.default_rng(seed)
  W = 128
  fov_up_deg, fov_down_deg = fov
  depth_im, color_im = _scene(rng, H, W, fov_up_deg, fov_down_deg, zero_frac)
  dx, dy, dz = 36, 32, 24
  org = (-(rng.uniform(0.2, 0.8) * np.array([dx, dy, dz]) * vox)).astype(F)
  trunc = F(5 * vox)
  fov_up, fov_down = F(fov_up_deg * np.pi / 180.0), F(fov_down_deg * np.pi / 180.0)
  vx, vy, vz = np.meshgrid(np.arange(dx, dtype=F), np.arange(dy, dtype=F), np.arange(dz, dtype=F), indexing="ij")
  x, y, z = _fma(vx, F(vox), org[0]), _fma(vy, F(vox), org[1]), _fma(vz, F(vox), org[2])
  yaw = (-np.arctan2(y, x)).astype(F)
  px = np.clip(np.floor((0.5 * (yaw.astype(np.float64) / np.pi + 1.0)).astype(F) * F(W)), 0, W - 1).astype(int)
  changes = _reference_changes(rng, x, y, z, px, depth_im, color_im, H, fov_up, fov_down, trunc)
  xm = max(abs(float(org[0])), abs(float(org[0]) + dx * vox))
  ym = max(abs(float(org[1])), abs(float(org[1]) + dy * vox))
  zlo, zhi = _hulls(x[:, :, 0], y[:, :, 0], px[:, :, 0], depth_im, color_im, H, W, fov_up, fov_down, trunc, float(org[2]), vox, dz,
                    float(np.hypot(xm, ym)), mutate)
  k = np.arange(dz)[None, None, :]
  inside = (k >= zlo[:, :, None]) & (k <= zhi[:, :, None])
  return changes, inside


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), H=st.sampled_from([16, 32, 64, 128]), fov=st.sampled_from([(3.0, -25.0), (10.67, -30.67),
       (22.5, -22.5), (2.0, -24.8), (15.0, -15.0), (34.0, -34.0)]), vox=st.sampled_from([0.05, 0.1, 0.25, 0.4]),
       zero_frac=st.sampled_from([0.0, 0.0, 0.3]))
def test_hull_contains_every_voxel_the_kernel_string_would_change(seed, H, fov, vox, zero_frac):
  changes, inside = _case(seed, H, fov, vox, zero_frac)
  bad = changes & ~inside
  assert not bad.any(), (int(bad.sum()), np.argwhere(bad)[:3].tolist())
  if zero_frac == 0.0 and changes.any():
    assert inside.mean() < 0.6, inside.mean()       # the hulls are worth having


def test_the_margins_matter():
  """Each mutation of a margin lets at least one changed voxel fall outside its hull on some seed: the test has teeth.
  (The extra voxel at each end of a hull and the extra bin at each end of a range-table entry, `no_voxel_slack` /
  `no_bin_slack`, are belt and braces on top of floor / ceil and the 1 mm offsets: none of these cases needs them.)"""
  for mutate in ("no_pitch_slack", "half_shell"):
    caught = False
    for seed in range(40):
      for H, fov, vox in ((64, (3.0, -25.0), 0.1), (128, (22.5, -22.5), 0.05), (16, (10.67, -30.67), 0.25)):
        changes, inside = _case(seed, H, fov, vox, 0.0, mutate)
        if (changes & ~inside).any():
          caught = True
          break
      if caught:
        break
    assert caught, mutate
