"""CPU properties of the iso-surface extraction restatement (oracle.mesh_extract) and of the derived
marching-cubes table (tools/gen_mc_table.py): closed, consistently oriented surfaces at the right place."""
import numpy as np
import pytest


def _sdf_volume(shape, fn):
  g = np.stack(np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij"), -1)
  return fn(g).astype(np.float32)


def _edges_closed(verts, faces):
  """Every undirected edge of a closed surface is used by exactly two triangles, once per direction."""
  uniq, inv = np.unique(verts.round(5), axis=0, return_inverse=True)
  f = inv.reshape(-1)[faces]
  f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]  # drop zero-area triangles
  d = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
  und = np.sort(d, axis=1)
  _, counts = np.unique(und, axis=0, return_counts=True)
  directed, dcounts = np.unique(d, axis=0, return_counts=True)
  return counts, dcounts, len(uniq), len(f)


@pytest.mark.parametrize("radius,centre", [(9.3, (15.2, 14.7, 16.1)), (5.05, (8.5, 9.0, 7.5))])
def test_sphere_is_closed_oriented_and_on_the_surface(oracle, radius, centre):
  c = np.array(centre)
  vol = _sdf_volume((32, 30, 33), lambda g: (np.linalg.norm(g - c, axis=-1) - radius) / 3.0)
  z = np.zeros_like(vol)
  m = oracle.mesh_extract(vol, z, z, 1.0, np.zeros(3, np.float32))
  assert m["faces"].shape[0] > 200
  assert np.array_equal(m["faces"].reshape(-1), np.arange(3 * m["faces"].shape[0]))
  r = np.linalg.norm(m["verts"].astype(np.float64) - c, axis=1)
  assert np.abs(r - radius).max() < 0.15  # linear interpolation error of a curved SDF
  counts, dcounts, n_v, n_f = _edges_closed(m["verts"], m["faces"])
  assert (counts == 2).all() and (dcounts == 1).all()
  assert n_v - len(counts) + n_f == 2  # Euler characteristic of a sphere
  # outward orientation (inside = value < level): signed volume is +4/3 pi r^3
  v = m["verts"].astype(np.float64)[m["faces"]]
  signed = np.einsum("ij,ij->i", v[:, 0] - c, np.cross(v[:, 1] - c, v[:, 2] - c)).sum() / 6.0
  assert abs(signed / (4.0 / 3.0 * np.pi * radius ** 3) - 1.0) < 0.03


def test_random_volume_is_watertight(oracle):
  """All 256 cases incl. the ambiguous ones: a random field still gives a closed 2-manifold-edge surface
  (interior edges used exactly twice; edges on the volume border are open)."""
  rng = np.random.default_rng(5)
  vol = rng.uniform(-1, 1, (14, 13, 12)).astype(np.float32)
  vol[[0, -1], :, :] = 1.0; vol[:, [0, -1], :] = 1.0; vol[:, :, [0, -1]] = 1.0  # close the border
  z = np.zeros_like(vol)
  m = oracle.mesh_extract(vol, z, z, 1.0, np.zeros(3, np.float32))
  counts, dcounts, _, n_f = _edges_closed(m["verts"], m["faces"])
  assert n_f > 1000
  assert (counts % 2 == 0).all()  # ambiguous-face junctions may stack two sheets on one edge, never leave it open
  assert (dcounts <= 2).all()


def test_plane_world_transform_and_attribute_lookup(oracle):
  """A tilted plane: vertices satisfy the plane equation in WORLD coordinates (verts * voxel + origin,
  fusion_lidar.py:412) and colours / remissions come from the nearest voxel with the uint8 wrap (:409-423)."""
  n = np.array([0.2, -0.1, 1.0]); n /= np.linalg.norm(n)
  vol = _sdf_volume((20, 18, 16), lambda g: (g @ n - 7.3) / 2.5)
  rng = np.random.default_rng(1)
  labels = rng.choice([40, 48, 70, 259, 300], size=vol.shape).astype(np.float32)
  color_vol = labels * 65536.0
  rem_vol = rng.random(vol.shape).astype(np.float32)
  vox, origin = 0.25, np.array([-3.0, 2.0, -1.5], np.float32)
  m = oracle.mesh_extract(vol, color_vol, rem_vol, vox, origin)
  vv = (m["verts"].astype(np.float64) - origin) / vox
  assert np.abs(vv @ n - 7.3).max() < 1e-4
  att_v, att_c, att_r = oracle.mesh_attributes(vv.astype(np.float32), color_vol, rem_vol, vox, origin)
  near = np.abs(vv - np.round(vv)).max(axis=1) < 0.499  # away from rounding ties of the float32 round trip
  assert near.mean() > 0.9
  assert np.array_equal(m["colors"][near], att_c[near])
  assert np.array_equal(m["rem"][near], att_r[near])
  assert set(np.unique(m["colors"][:, 2])) <= {40, 48, 70, 259 - 256, 300 - 256}
  assert (m["colors"][:, :2] == 0).all()
