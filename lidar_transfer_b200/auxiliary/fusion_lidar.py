"""Drop-in for the reference's `auxiliary.fusion_lidar` (auxiliary/fusion_lidar.py): same class,
method names, argument order and return tuples; the four volumes live in HBM as torch tensors and
every stage runs through libvlidar (include/vlidar.h):

  TSDFVolume.__init__            fusion_lidar.py:23-63    -> vl_tsdf_init (no host volumes, no upload)
  TSDFVolume.integrate           fusion_lidar.py:252-287  -> vl_tsdf_integrate
  TSDFVolume.get_volume          fusion_lidar.py:395-400  -> D2H on request only
  TSDFVolume.get_mesh            fusion_lidar.py:403-424  -> vl_mesh_extract (iso-surface + vertex lookup on device)
  TSDFVolume.throw_rays_at_mesh  fusion_lidar.py:426-455  -> vl_beams_build + vl_cast on the device-resident mesh
                                                           (vl_bvh_build + vl_trace if the cast reports VL_ENOSPACE)
  meshwrite                      fusion_lidar.py:462-495  -> same ASCII PLY bytes, vectorised

The host-facing values (numpy arrays, dtypes, shapes) are the reference's; `*_device` variants return
CUDA tensors without the D2H copies for callers that stay on the GPU.
"""
import numpy as np
import torch

from .. import _lib, engine

class LazyHostArray(object):
  """A device tensor standing in for a numpy array the reference returns but its batch driver never reads (the
  mesh of deform(): 100-250 MB per scan that only the GUI and the optional PLY dump look at).  Shape / dtype / len are
  known at once; the device -> host copy happens on first real use (np.asarray, indexing, arithmetic) and is kept."""

  def __init__(self, tensor, fn=None):
    self._t, self._fn, self._a = tensor, fn, None
    shape = tuple(tensor.shape) if fn is None else None
    if fn is not None:            # fn maps the tensor to the final device tensor lazily (e.g. a colour look-up)
      self._shape, self._dtype = fn("shape"), fn("dtype")
    else:
      self._shape, self._dtype = shape, np.dtype(str(tensor.dtype).replace("torch.", ""))

  shape = property(lambda self: self._shape)
  dtype = property(lambda self: self._dtype)
  ndim = property(lambda self: len(self._shape))
  size = property(lambda self: int(np.prod(self._shape)))

  def __len__(self):
    return self._shape[0]

  def device_tensor(self):
    return self._t if self._fn is None else self._fn(self._t)

  def __array__(self, dtype=None, copy=None):
    if self._a is None:
      self._a = self.device_tensor().cpu().numpy()
    return self._a if dtype is None else self._a.astype(dtype, copy=False)

  def __getitem__(self, key):
    return self.__array__()[key]

  def __mul__(self, other):
    return self.__array__() * other

  __rmul__ = __mul__

  def __repr__(self):
    return "LazyHostArray(shape=%s, dtype=%s, %s)" % (self._shape, self._dtype, "host copy made" if self._a is not None else "on device")


FUSION_GPU_MODE = 1  # the reference sets 0 when pycuda is missing and falls back to numpy; there is no fallback here


class TSDFVolume(object):

  def __init__(self, vol_bnds, voxel_size, fov_up, fov_down):
    engine.require_cuda()
    # Define projection parameters
    self.fov_up = fov_up
    self.fov_down = fov_down

    # Define voxel volume parameters (fusion_lidar.py:29-38; vol_bnds is adjusted IN PLACE like the reference)
    self._vol_bnds = vol_bnds
    self._voxel_size = voxel_size
    self._trunc_margin = self._voxel_size * 5
    self._vol_dim = np.ceil((self._vol_bnds[:, 1] - self._vol_bnds[:, 0]) /
                            self._voxel_size).copy(order='C').astype(int)
    self._vol_bnds[:, 1] = self._vol_bnds[:, 0] + self._vol_dim * self._voxel_size
    self._vol_origin = self._vol_bnds[:, 0].copy(order='C').astype(np.float32)
    print("Voxel volume size: %d x %d x %d" % (self._vol_dim[0], self._vol_dim[1], self._vol_dim[2]))
    print("Voxel volume [m]: %d x %d x %d" % (self._vol_dim[0] * voxel_size,
                                              self._vol_dim[1] * voxel_size,
                                              self._vol_dim[2] * voxel_size))
    print("Voxel count: %d mio" % (self._vol_dim[0] * self._vol_dim[1] * self._vol_dim[2] / 1E6))
    print("Voxel size: %f" % (self._voxel_size))
    self._dev = engine.TsdfDevice(self._vol_dim, self._vol_origin, self._voxel_size, self.fov_up, self.fov_down)
    self._mesh = None

  def integrate(self, color_im, depth_im, rem_im, cam_pose, obs_weight=1.):
    """ Data should be in world frame with pose transformation applied
        Not using the cam_pose input!  (fusion_lidar.py:252-287)
    """
    # Fold RGB color image into a single channel image (fusion_lidar.py:259-264)
    if torch.is_tensor(color_im):
      c = color_im.to(torch.float32)
      color_im = torch.floor(c[:, :, 0] * 256 * 256 + c[:, :, 1] * 256 + c[:, :, 2])
    else:
      color_im = np.asarray(color_im).astype(np.float32)
      color_im = np.floor(color_im[:, :, 0] * 256 * 256 + color_im[:, :, 1] * 256 + color_im[:, :, 2])
    self._dev.integrate(color_im, depth_im, rem_im, obs_weight)
    self._mesh = None

  def integrate_device(self, label_im, depth_im, rem_im, obs_weight=1.):
    """integrate() for images that are already on the device: label_im i32[H,W] is channel 0 of the reference's
    colour image (proj_label3[:, :, 0] = label, the other two zero: laserscan.py:970-973), whose fold :259-264 is
    label * 65536 -- exact in float32 for every 16-bit label."""
    self._dev.integrate(label_im.to(torch.float32) * 65536.0, depth_im, rem_im, obs_weight)
    self._mesh = None

  # Copy voxel volume to CPU
  def get_volume(self):
    return self._dev.tsdf.cpu().numpy(), self._dev.color.cpu().numpy(), self._dev.rem.cpu().numpy()

  def get_mesh_device(self, want_norms=False):
    """Device-resident mesh: dict(verts f32[N_v,3], faces i32[N_t,3], norms f32[N_v,3] or None, colors u8[N_v,3],
    rem f32[N_v]).  The normals are only written when asked for (the ray cast does not read them)."""
    if self._mesh is None or (want_norms and self._mesh["norms"] is None):
      self._mesh = self._dev.extract_mesh(want_norms=want_norms)
    return self._mesh

  # Get mesh of voxel volume via marching cubes
  def get_mesh(self, color_lut):
    m = self.get_mesh_device(want_norms=True)
    return (m["verts"].cpu().numpy(), m["faces"].cpu().numpy(), m["norms"].cpu().numpy(),
            m["colors"].cpu().numpy(), m["rem"].cpu().numpy())

  _beams_cache = {}   # (ray buffer id, H, device) -> (rays, Beams): the beam index is the target sensor's constant

  def throw_rays_at_mesh_device(self, rays, origin, H, W):
    m = self.get_mesh_device()
    colors = m["colors"]   # uint8, read as it is by the cast (the reference converts: colors.astype(np.int32), :437)
    dev = m["verts"].device
    R = int(np.asarray(rays).size // 3) if not torch.is_tensor(rays) else rays.numel() // 3
    # the four per-ray outputs as views of ONE buffer (a single device -> host copy for the caller)
    packed = torch.empty(32 * R, dtype=torch.uint8, device=dev)
    outs = dict(packed=packed, endpoints=packed[:12 * R].view(torch.float32), endcolors=packed[12 * R:24 * R].view(torch.int32),
                range=packed[24 * R:28 * R].view(torch.float32), endrem=packed[28 * R:32 * R].view(torch.float32))
    try:
      key = (id(rays), int(H), str(dev))
      hit = TSDFVolume._beams_cache.get(key)
      if hit is None or hit[0] is not rays:
        if len(TSDFVolume._beams_cache) > 8:
          TSDFVolume._beams_cache.clear()
        hit = TSDFVolume._beams_cache[key] = (rays, engine.Beams(rays, H, dev))
      beams = hit[1]
      out = engine.cast(beams, m["verts"], m["faces"], colors, m["rem"], origin, out=outs, want_ids=False, zero_misses=True)
    except _lib.VlidarError as e:
      if e.code != _lib.VL_ENOSPACE:
        raise
      bvh = engine.Bvh(m["verts"], m["faces"], colors.to(torch.int32), m["rem"])
      out = engine.trace(bvh, rays, origin, H, out=outs, want_ids=False, zero_misses=True)
    return out, m

  def throw_rays_at_mesh(self, rays, origin, H, W, color_lut):
    print("Get mesh by marching cubes...")
    print("Raytracing...")
    out, m = self.throw_rays_at_mesh_device(rays, origin, H, W)
    # the per-ray results as numpy like the reference (one device -> host copy of the packed buffer); the mesh
    # (elements 2-4) lazily: see LazyHostArray
    a = out["packed"].cpu().numpy()
    R = a.size // 32
    return a[:12 * R].view(np.float32).reshape(-1, 3), a[12 * R:24 * R].view(np.int32).reshape(-1, 3), \
        LazyHostArray(m["verts"]), LazyHostArray(m["colors"]), LazyHostArray(m["faces"]), \
        a[24 * R:28 * R].view(np.float32).reshape(-1, W), a[28 * R:32 * R].view(np.float32).reshape(-1, W)


# ------------------------------------------------------------------------------
# Additional helper functions

# Save 3D mesh to a polygon .ply file  (fusion_lidar.py:462-495: same header, same "%f"/"%d" rows)
def meshwrite(filename, verts, faces, norms, colors):
  verts, faces, norms, colors = np.asarray(verts), np.asarray(faces), np.asarray(norms), np.asarray(colors)
  with open(filename, 'w') as ply_file:
    ply_file.write("ply\n")
    ply_file.write("format ascii 1.0\n")
    ply_file.write("element vertex %d\n" % (verts.shape[0]))
    ply_file.write("property float x\n")
    ply_file.write("property float y\n")
    ply_file.write("property float z\n")
    ply_file.write("property float nx\n")
    ply_file.write("property float ny\n")
    ply_file.write("property float nz\n")
    ply_file.write("property uchar red\n")
    ply_file.write("property uchar green\n")
    ply_file.write("property uchar blue\n")
    ply_file.write("element face %d\n" % (faces.shape[0]))
    ply_file.write("property list uchar int vertex_index\n")
    ply_file.write("end_header\n")
    if verts.shape[0]:
      table = np.concatenate([verts[:, :3].astype(np.float64), norms[:, :3].astype(np.float64)], axis=1)
      cols = np.trunc(colors[:, :3].astype(np.float64)).astype(np.int64)  # "%d" truncates floats towards zero
      fmt = "%f %f %f %f %f %f %d %d %d\n"
      chunk = 65536
      for s in range(0, verts.shape[0], chunk):
        t, c = table[s:s + chunk], cols[s:s + chunk]
        ply_file.write("".join(fmt % (r[0], r[1], r[2], r[3], r[4], r[5], k[0], k[1], k[2])
                               for r, k in zip(t.tolist(), c.tolist())))
    if faces.shape[0]:
      f = faces.astype(np.int64)
      for s in range(0, f.shape[0], 65536):
        ply_file.write("".join("3 %d %d %d\n" % (a, b, c) for a, b, c in f[s:s + 65536].tolist()))
