"""Seeded synthetic "KITTI-shape" scenes and sensors for benchmarks and parity tests.

There is no network for datasets, so BASELINE.json configs 3 and 5 use this generator
(SURVEY.md section 8d): a ground height field over x,y in [-50,50] m with a radial berm so
every ray terminates, plus axis-aligned boxes (cars, buildings; none of them over the sensor)
tessellated at 4x the grid pitch.  Per-vertex labels are drawn from the SemanticKITTI ids the reference's config
lists (config/lidar_transfer.yaml:13-46) excluding 0, 1 and ids >= 256; remission U[0,1).

The mesh layout is the one TSDFVolume.get_mesh hands to C_Trace
(auxiliary/fusion_lidar.py:434-438): verts float32[N_v,3], faces int32[N_t,3],
colors int32[N_v,3] = (0, 0, label), rem float32[N_v].
"""
import numpy as np

# SemanticKITTI ids of config/lidar_transfer.yaml:13-46 minus {0, 1} and the moving classes >= 252
STATIC_LABELS = np.array([10, 11, 13, 15, 16, 18, 20, 30, 31, 32, 40, 44, 48, 49, 50, 51, 52,
                          60, 70, 71, 72, 80, 81, 99], np.int32)

# sensors: name -> (beams H, columns W, fov_up, fov_down); W = int(fov_hor / angle_res_hor)
SENSORS = {
    "HDL-64E": (64, 2048, 3.0, -25.0),       # minimal/config.yaml
    "HDL-32E": (32, 1024, 10.67, -30.67),    # minimal/target.yaml
    "OS1-128": (128, 2048, 22.5, -22.5),     # BASELINE.json config 4
}


def grid_side_for_triangles(n_tris):
  """Ground grid side n so that 2 (n-1)^2 is closest to n_tris (500 -> 498 002, 710 -> 1 005 362)."""
  return max(2, int(round(np.sqrt(n_tris / 2.0))) + 1)


def _patch(corner, eu, ev, nu, nv):
  """(nu+1)x(nv+1) vertex grid spanning corner + s*eu + t*ev, two triangles per cell."""
  s = np.linspace(0.0, 1.0, nu + 1)
  t = np.linspace(0.0, 1.0, nv + 1)
  S, T = np.meshgrid(s, t, indexing="ij")
  v = corner[None, None, :] + S[..., None] * eu[None, None, :] + T[..., None] * ev[None, None, :]
  idx = np.arange((nu + 1) * (nv + 1)).reshape(nu + 1, nv + 1)
  a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
  f = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
  return v.reshape(-1, 3), f


def make_scene(seed, n_side=500, n_boxes=40, extent=50.0):
  """Returns dict(verts f32[N_v,3], faces i32[N_t,3], colors i32[N_v,3], rem f32[N_v], labels i32[N_v])."""
  rng = np.random.default_rng(seed)
  xs = np.linspace(-extent, extent, n_side)
  X, Y = np.meshgrid(xs, xs, indexing="ij")
  Z = -1.73 + 0.3 * np.sin(0.3 * X) * np.cos(0.2 * Y) + rng.normal(0.0, 0.05, X.shape)
  R = np.hypot(X, Y)
  Z = Z + np.where(R > 40.0, 1.5 * (R - 40.0), 0.0)
  verts = [np.stack([X, Y, Z], -1).reshape(-1, 3)]
  idx = np.arange(n_side * n_side).reshape(n_side, n_side)
  a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
  faces = [np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])]
  n_v = verts[0].shape[0]
  pitch = 2.0 * extent / (n_side - 1)
  for k in range(n_boxes):
    if k % 4 == 0:   # building
      sx, sy, sz = rng.uniform(10.0, 20.0), 10.0, 8.0
    else:            # car
      sx, sy, sz = 4.0, 1.8, 1.5
    if rng.random() < 0.5:
      sx, sy = sy, sx
    while True:  # keep the sensor (origin) outside every box, with 2.5 m clearance
      r = rng.uniform(6.0, 38.0)
      th = rng.uniform(0.0, 2.0 * np.pi)
      cx, cy, z0 = r * np.cos(th), r * np.sin(th), -2.2
      if abs(cx) > sx / 2 + 2.5 or abs(cy) > sy / 2 + 2.5:
        break
    lo = np.array([cx - sx / 2, cy - sy / 2, z0])
    ex, ey, ez = np.array([sx, 0, 0.0]), np.array([0, sy, 0.0]), np.array([0, 0, sz + 0.5])
    # boxes are tessellated at 4x the ground pitch: 40 boxes add ~4 % to the triangle budget
    nx, ny, nz = (max(1, int(round(s / (4.0 * pitch)))) for s in (sx, sy, sz + 0.5))
    for corner, eu, ev, nu, nv in ((lo, ex, ez, nx, nz), (lo + ey, ex, ez, nx, nz),
                                   (lo, ey, ez, ny, nz), (lo + ex, ey, ez, ny, nz),
                                   (lo + ez, ex, ey, nx, ny)):
      v, f = _patch(corner, eu, ev, nu, nv)
      verts.append(v)
      faces.append(f + n_v)
      n_v += v.shape[0]
  verts = np.ascontiguousarray(np.concatenate(verts).astype(np.float32))
  faces = np.ascontiguousarray(np.concatenate(faces).astype(np.int32))
  labels = rng.choice(STATIC_LABELS, size=verts.shape[0]).astype(np.int32)
  colors = np.zeros((verts.shape[0], 3), np.int32)
  colors[:, 2] = labels
  rem = rng.random(verts.shape[0]).astype(np.float32)
  return dict(verts=verts, faces=faces, colors=colors, rem=rem, labels=labels)


def make_scan_points(seed, n_points=124668, fov_up=3.0, fov_down=-25.0):
  """A KITTI-shaped point cloud (float32[N,4] x,y,z,remission + uint32 labels) sampled from the
  analytic ground/berm surface plus noise, for the projection / TSDF kernels."""
  rng = np.random.default_rng(seed)
  yaw = rng.uniform(-np.pi, np.pi, n_points)
  pitch = np.deg2rad(rng.uniform(fov_down, fov_up, n_points))
  # range to the ground plane z = -1.73 along the beam, capped to the berm radius
  with np.errstate(divide="ignore"):
    rng_ground = np.where(pitch < -0.01, -1.73 / np.sin(np.minimum(pitch, -0.01)), 80.0)
  r = np.minimum(rng_ground, rng.uniform(30.0, 80.0, n_points)) * (1.0 + rng.normal(0, 0.002, n_points))
  x = r * np.cos(pitch) * np.cos(yaw)
  y = r * np.cos(pitch) * np.sin(yaw)
  z = r * np.sin(pitch)
  pts = np.stack([x, y, z, rng.random(n_points)], -1).astype(np.float32)
  labels = rng.choice(STATIC_LABELS, size=n_points).astype(np.uint32)
  return pts, labels
