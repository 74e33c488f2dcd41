"""The arithmetic behind the TSDF shell sweep (lidar_transfer_b200/csrc/vl_tsdf.cu: shell_bracket, k_tsdf_shell),
restated in numpy float32 and checked on the CPU over many sensor / volume configurations: the conservative bracket
(plain reciprocal square root, arcsine series to s^15, one or two candidate image rows) must NEVER rule out a voxel
that the kernel string's own decision (fusion_lidar.py:119-227 on a fresh volume) would change -- whatever the last
bits of norm3df / asinf are -- and it must rule out most of the others.  The GPU tests compare the kernels bit for bit
on nine configurations; this test covers the space between them."""
import numpy as np
from hypothesis import given, settings, strategies as st

F = np.float32
K_ASIN_ERR, K_DEPTH_REL = 1e-5, 1e-5          # vl_tsdf.cu: kAsinErr, kDepthRel


def _fma(a, b, c):
  return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(F)


def _scene(rng, H, W, fov_up, fov_down, zero_frac):
  """A range / label image of a ground plane plus random boxes, with empty and label-0 pixels."""
  pitch = np.deg2rad(np.linspace(fov_up, fov_down, H))[:, None] * np.ones((1, W))
  depth = np.where(pitch < -0.02, 1.73 / np.maximum(np.sin(-pitch), 1e-3), 0.0)
  depth = np.minimum(depth, 60.0) * (pitch < -0.02)
  for _ in range(6):
    c0, w, d = rng.integers(0, W), rng.integers(3, W // 3 + 4), rng.uniform(2.0, 25.0)
    cols = (c0 + np.arange(w)) % W
    r0 = rng.integers(0, H)
    blk = depth[r0:, cols]
    depth[r0:, cols] = np.where((blk == 0) | (blk > d), d, blk)
  depth = (depth * (1 + 0.01 * rng.standard_normal(depth.shape))).astype(F)
  depth[rng.random(depth.shape) < 0.05] = 0.0
  label = rng.integers(2, 250, depth.shape)
  label[rng.random(depth.shape) < zero_frac] = 0
  return depth, (label * 65536).astype(F)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), H=st.sampled_from([16, 32, 64, 128, 512]), fov=st.sampled_from([(3.0, -25.0), (10.67, -30.67),
       (22.5, -22.5), (2.0, -24.8), (15.0, -15.0), (34.0, -34.0)]), vox=st.sampled_from([0.05, 0.1, 0.25, 0.4]),
       zero_frac=st.sampled_from([0.0, 0.0, 0.3]))
def test_bracket_never_rules_out_a_voxel_the_kernel_string_would_change(seed, H, fov, vox, zero_frac):
  rng = np.random.default_rng(seed)
  W = 256
  fov_up_deg, fov_down_deg = fov
  depth_im, color_im = _scene(rng, H, W, fov_up_deg, fov_down_deg, zero_frac)
  dx, dy, dz = 40, 36, 28
  org = (-(rng.uniform(0.2, 0.8) * np.array([dx, dy, dz]) * vox)).astype(F)      # the sensor is inside the volume
  trunc = F(5 * vox)
  fov_up, fov_down = F(fov_up_deg * np.pi / 180.0), F(fov_down_deg * np.pi / 180.0)
  fov_tot = F(abs(fov_up) + abs(fov_down))
  vx, vy, vz = np.meshgrid(np.arange(dx, dtype=F), np.arange(dy, dtype=F), np.arange(dz, dtype=F), indexing="ij")
  x, y, z = _fma(vx, F(vox), org[0]), _fma(vy, F(vox), org[1]), _fma(vz, F(vox), org[2])
  # image column: depends on x and y only -- the kernel reads it from the per-column table, both sides share it here
  yaw = (-np.arctan2(y, x)).astype(F)
  px = np.clip(np.floor((0.5 * (yaw.astype(np.float64) / np.pi + 1.0)).astype(F) * F(W)), 0, W - 1).astype(int)

  # ---- the kernel string's decision.  Its norm3df / asinf are not glibc's to the last bit; instead of guessing those
  # bits, depth and pitch are pushed per voxel by up to 0.9 of what the bracket claims to tolerate (kDepthRel / 2 relative,
  # kAsinErr absolute): the bracket has to hold for ANY such implementation ------------------------------------------
  with np.errstate(invalid="ignore", divide="ignore"):
    d64 = np.sqrt(x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2 + z.astype(np.float64) ** 2)
    depth = (d64 * (1 + 0.45 * K_DEPTH_REL * rng.uniform(-1, 1, d64.shape))).astype(F)
    pitch = (np.arcsin(np.clip(z.astype(np.float64) / d64, -1, 1)) + 0.9 * K_ASIN_ERR * rng.uniform(-1, 1, d64.shape)).astype(F)
    in_fov = ~((pitch > fov_up) | (pitch < fov_down))                    # NaN (origin voxel) passes, like the reference
    proj_y = (1.0 - (pitch.astype(np.float64) + abs(np.float64(fov_down))) / np.float64(fov_tot)).astype(F) * F(H)
    row = np.clip(np.nan_to_num(np.floor(proj_y), nan=0.0), 0, H - 1).astype(int)
    dv = depth_im[row, px]
    diff = dv - depth
    changes = in_fov & (dv != 0) & ~(diff < -trunc) & ((color_im[row, px] == 0) | (np.minimum(F(1), diff / trunc) < 0))

  # ---- the bracket (shell_bracket + k_tsdf_shell) ----------------------------------------------------------------------
  lo = np.where(depth_im == 0, np.inf, np.where(color_im == 0, -np.inf, depth_im)).astype(F)
  hi = np.where(depth_im == 0, -np.inf, depth_im + trunc).astype(F)
  h_over_fov = F(H) / fov_tot
  eps_row = F(1.02 * K_ASIN_ERR * H / float(fov_tot) + 2e-4 + 4e-7 * H)
  assert eps_row < 0.45
  with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
    q = _fma(z, z, _fma(x, x, (y * y).astype(F)))
    rq = (1.0 / np.sqrt(q.astype(np.float64))).astype(F)
    rq = rq * F(1 + rng.choice([-2, 0, 2]) * 2.0 ** -23)                 # rsqrt.approx: 2 ulp
    d = q * rq
    sn = z * rq
    s2 = sn * sn
    poly = F(135135 / 9676800)
    for c in (10395 / 599040, 945 / 42240, 105 / 3456, 15 / 336, 3 / 40, 1 / 6):
      poly = _fma(s2, poly, F(c))
    pit = _fma(sn * s2, poly, sn)
    outside = (pit > fov_up + F(K_ASIN_ERR)) | (pit < fov_down - F(K_ASIN_ERR))
    rowf = _fma(-(pit + abs(fov_down)), h_over_fov, F(H))
    f0 = np.floor(rowf - eps_row)
    r0 = np.nan_to_num(f0, nan=0.0).astype(int)
    r1 = np.where(rowf + eps_row >= f0 + 1, r0 + 1, r0)
    r0, r1 = np.clip(r0, 0, H - 1), np.clip(r1, 0, H - 1)
    d_lo, d_hi = d * F(1 - K_DEPTH_REL), d * F(1 + K_DEPTH_REL)
    may0 = ~((d_hi < lo[r0, px]) | (d_lo > hi[r0, px]))
    may1 = ~((d_hi < lo[r1, px]) | (d_lo > hi[r1, px]))
    exact = ~outside & (may0 | ((r1 != r0) & may1))
  skipped = ~exact
  bad = changes & skipped
  assert not bad.any(), (int(bad.sum()), np.argwhere(bad)[:3].tolist())
  if abs(fov_up_deg) <= 35 and abs(fov_down_deg) <= 35:
    assert changes.sum() <= exact.sum()
    if zero_frac == 0.0:
      assert skipped.mean() > 0.6, skipped.mean()      # the bracket is worth having: most voxels never reach the arithmetic


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), dy=st.integers(1, 4096), dz=st.integers(1, 4096))
def test_float_index_decode_equals_the_integer_decode_away_from_slab_boundaries(seed, dy, dz):
  """The kernel string decodes voxel_idx with float32 divisions (fusion_lidar.py:96-98), which lands in the neighbouring
  slab for some indices beyond 2^24.  The shell sweep uses the integer decode for voxels at least kDecodeWindow = 256
  away from a slab boundary (and dy * dz <= 2^24) and the float decode for the rest; the two must agree there, for any
  volume below 2^31 voxels."""
  rng = np.random.default_rng(seed)
  slab = dy * dz
  if slab > (1 << 24) or slab <= 512:
    return
  dx = int(min(65535, ((1 << 31) - 1) // slab))
  n = 20000
  x = rng.integers(0, dx, n)
  rem = np.concatenate([rng.integers(256, slab - 256, n - 4), [256, 257, slab - 257, slab - 258]])
  x[-4:] = dx - 1                                                  # the largest indices: the coarsest float spacing
  idx = (x * slab + rem).astype(np.int64)
  assert idx.max() < (1 << 31)
  fx = np.floor(idx.astype(np.float32) / np.float32(slab))
  r1 = (idx - fx.astype(np.int64) * slab)
  fy = np.floor(r1.astype(np.float32) / np.float32(dz))
  fz = r1 - fy.astype(np.int64) * dz
  assert np.array_equal(fx.astype(np.int64), x)
  assert np.array_equal(fy.astype(np.int64), rem // dz) and np.array_equal(fz, rem % dz)
  # the sweep's own y / z decode: one multiplication by 1 / dz and a correction by one
  vy = (rem.astype(np.float32) * (np.float32(1.0) / np.float32(dz))).astype(np.int64)
  vz = rem - vy * dz
  vy = np.where(vz < 0, vy - 1, np.where(vz >= dz, vy + 1, vy))
  vz = rem - vy * dz
  assert np.array_equal(vy, rem // dz) and (vz >= 0).all() and (vz < dz).all()
