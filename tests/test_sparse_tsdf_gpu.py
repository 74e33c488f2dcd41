"""Sparse (blocked) TSDF volumes -- the reference's own TODO, auxiliary/fusion_lidar.py:45 -- against the dense path,
bit for bit: vl_tsdf_sparse_integrate keeps, per z column, only an interval of voxels (the hull); everything outside
is the initial state by definition.  After vl_tsdf_densify the four arrays must equal vl_tsdf_init + vl_tsdf_integrate
voxel for voxel; the mesh extracted straight from the sparse volume (vl_mesh_*_sparse) must equal the dense mesh,
vertex for vertex and in the same order.  (Against the reference's own CUDA kernel: tests/test_reference_driver_gpu.py,
which runs with sparse volumes, the default.)"""
import numpy as np
import pytest
import torch

from lidar_transfer_b200 import synth

pytestmark = pytest.mark.gpu


def _images(oracle, seed, fov, H, W, zero_label_frac=0.0, junk=False, n=124668):
  pts, labels = synth.make_scan_points(seed, n, fov_up=fov[0], fov_down=fov[1])
  pr = oracle.project(pts[:, :3].astype(np.float64), pts[:, 3], labels, fov[0], fov[1], H, W)
  color = oracle.label_to_color_im(pr["proj_label"])
  depth, rem = pr["range_image"].copy(), pr["proj_remissions"].copy()
  rng = np.random.default_rng(seed)
  if zero_label_frac:     # label-0 pixels: the running-average branch fires anywhere in front of them (fusion_lidar.py:200-211)
    color[rng.random(color.shape) < zero_label_frac] = 0.0
  if junk:                # NaN / inf / negative depths
    m = rng.random(depth.shape)
    depth[m < 0.001] = np.nan
    depth[(m > 0.001) & (m < 0.002)] = np.inf
    depth[(m > 0.002) & (m < 0.003)] = -3.0
  return color, depth, rem


CASES = {
    # name: (dims, origin, voxel, fov, image H x W, scans, zero-label fraction, junk depths)
    "c1-half-res": ((1000, 710, 50), (-50, -35.5, -3), 0.1, (3.0, -25.0), (64, 2048), 1, 0.0, False),
    "three-scans": ((500, 400, 40), (-25, -20, -3), 0.1, (3.0, -25.0), (64, 2048), 3, 0.0, False),
    "label0-and-junk": ((320, 256, 48), (-16, -12.8, -3), 0.1, (3.0, -25.0), (64, 1024), 2, 0.01, True),
    "os1-odd-dz": ((333, 257, 37), (-33.3, -25.7, -3.5), 0.2, (22.5, -22.5), (128, 1024), 2, 0.0, False),
    "hdl32": ((400, 400, 60), (-20, -20, -4), 0.1, (10.67, -30.67), (32, 1024), 2, 0.0, False),
    "sensor-inside-small": ((160, 160, 40), (-8.0, -8.0, -3.0), 0.1, (3.0, -25.0), (64, 2048), 2, 0.0, False),
}


@pytest.mark.parametrize("name", list(CASES))
def test_sparse_volume_equals_dense_volume_and_mesh(engine, oracle, name):
  dims, origin, vox, fov, (H, W), n_scans, zl, junk = CASES[name]
  origin = np.asarray(origin, np.float32)
  dense = engine.TsdfDevice(dims, origin, vox, fov[0], fov[1], sparse=False)
  sparse = engine.TsdfDevice(dims, origin, vox, fov[0], fov[1], sparse=True)
  n = dims[0] * dims[1] * dims[2]
  for k in range(n_scans):
    color, depth, rem = _images(oracle, 10 + k, fov, H, W, zl, junk)
    dense.integrate(color, depth, rem)
    sparse.integrate(color, depth, rem)
    # the mesh straight from the sparse volume, before anything densifies it
    ms = sparse.extract_mesh(want_norms=False)
    assert not sparse._dense
    md = dense.extract_mesh(want_norms=False)
    assert ms["faces"].shape[0] == md["faces"].shape[0]
    if k == 0 and not zl:
      assert md["faces"].shape[0] > 1000
    for key in ("verts", "colors", "rem"):
      a, b = ms[key], md[key]
      assert torch.equal(a.view(torch.uint8) if a.dtype == torch.uint8 else a.view(torch.int32),
                         b.view(torch.uint8) if b.dtype == torch.uint8 else b.view(torch.int32)), (name, k, key)
    # how sparse: voxels inside the hulls
    h = sparse._hull().cpu().numpy()
    lo, hi = h & 0xffff, (h >> 16) & 0xffff
    inside = int(np.maximum(hi - lo + 1, 0).sum())
    if name == "c1-half-res":
      assert inside < 0.15 * n, inside / n
    print("%s scan %d: %.2f %% of the voxels exist, %d triangles" % (name, k, 100.0 * inside / n, md["faces"].shape[0]))
  # dense views of the sparse volume
  for key in ("tsdf", "weight", "color", "rem"):
    a, b = getattr(sparse, key), getattr(dense, key)
    assert sparse._dense
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)), (name, key, int((a.view(torch.int32) != b.view(torch.int32)).sum()))
  # and a further integration on top of the densified volume still agrees
  color, depth, rem = _images(oracle, 99, fov, H, W, zl, junk)
  dense.integrate(color, depth, rem)
  sparse.integrate(color, depth, rem)
  assert torch.equal(sparse.tsdf.view(torch.int32), dense.tsdf.view(torch.int32))
  assert torch.equal(sparse.color.view(torch.int32), dense.color.view(torch.int32))


def test_sparse_limits_fall_back_to_the_dense_path(engine, oracle):
  """Outside the sweep's limits (field of view beyond 35 degrees here) vl_tsdf_sparse_integrate takes the dense path and
  marks every column whole: same volumes."""
  dims, origin, vox, fov = (96, 96, 40), np.asarray((-4.8, -4.8, -2.0), np.float32), 0.1, (40.0, -40.0)
  color, depth, rem = _images(oracle, 5, fov, 64, 512, n=40000)
  a = engine.TsdfDevice(dims, origin, vox, fov[0], fov[1], sparse=True)
  b = engine.TsdfDevice(dims, origin, vox, fov[0], fov[1], sparse=False)
  for _ in range(2):
    a.integrate(color, depth, rem)
    b.integrate(color, depth, rem)
  h = a._hull().cpu().numpy()
  assert ((h & 0xffff) == 0).all() and (((h >> 16) & 0xffff) == dims[2] - 1).all()
  for key in ("tsdf", "weight", "color", "rem"):
    assert torch.equal(getattr(a, key).view(torch.int32), getattr(b, key).view(torch.int32)), key
