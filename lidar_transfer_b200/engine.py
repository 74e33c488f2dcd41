"""Host side of the hot path: PyTorch tensors for device memory and streams, libvlidar for the work.

Everything here runs on the current CUDA device and the current torch stream; nothing
synchronises unless stated.  PyTorch is plumbing only -- no torch op computes any result.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def _stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
  return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _dev(x, dtype, device=None):
  """numpy / torch (any device) -> contiguous CUDA tensor of dtype."""
  if isinstance(x, np.ndarray):
    x = torch.from_numpy(np.ascontiguousarray(x))
  elif not torch.is_tensor(x):
    x = torch.as_tensor(x)
  if device is None:
    device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
  return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def require_cuda():
  if not torch.cuda.is_available():
    raise RuntimeError("lidar_transfer_b200 needs a CUDA device: the hot path has no CPU fallback")
  lib()


class Bvh:
  """(i) LBVH over an indexed triangle mesh, held in one device blob.

  Replaces Triangle construction + `BVH bvh(&objects)`, auxiliary/raytracer/RayTracer.cpp:32-54.
  verts f32[N_v,3], faces i32[N_t,3], colors i32[N_v,3], rem f32[N_v] -- the arrays
  TSDFVolume.throw_rays_at_mesh flattens for C_Trace (auxiliary/fusion_lidar.py:434-438).
  `blob` may be passed to reuse a previous allocation (no allocation in steady state).
  """

  def __init__(self, verts, faces, colors, rem, blob=None):
    require_cuda()
    self.verts = _dev(verts, torch.float32).reshape(-1)
    dev = self.verts.device
    self.faces = _dev(faces, torch.int32, dev).reshape(-1)
    self.colors = _dev(colors, torch.int32, dev).reshape(-1)
    self.rem = _dev(rem, torch.float32, dev).reshape(-1)
    self.n_verts = self.verts.numel() // 3
    self.n_faces = self.faces.numel() // 3
    if self.colors.numel() != 3 * self.n_verts or self.rem.numel() != self.n_verts:
      raise ValueError("colors must hold 3 ints and rem 1 float per vertex")
    need = lib().vl_bvh_blob_bytes(self.n_faces)
    if blob is None or blob.numel() < need or blob.device != dev:
      blob = torch.empty(need, dtype=torch.uint8, device=dev)
    self.blob = blob
    with torch.cuda.device(dev):
      check(lib().vl_bvh_build(_ptr(self.verts), _ptr(self.faces), _ptr(self.colors), _ptr(self.rem),
                               self.n_verts, self.n_faces, _ptr(self.blob), self.blob.numel(), _stream()))

  def status(self):
    """Synchronises; raises VlidarError(VL_EBADMESH) on out-of-range face indices.
    Returns dict(n_tris, root_ref, n_bad_faces, max_climb)."""
    info = (ctypes.c_int * 8)()
    with torch.cuda.device(self.blob.device):
      rc = lib().vl_bvh_status(_ptr(self.blob), self.n_faces, _stream(), info)
    out = dict(n_tris=info[0], root_ref=info[1], n_bad_faces=info[2], max_climb=info[3])
    check(rc)
    return out


TRACE_ZERO_MISSES = 1
RAYS_NORMALIZED = 4
COLORS_U8 = 8

# TSDF volumes are SPARSE by default (per z column only an interval of voxels exists, vl_tsdf_sparse_integrate); the
# properties tsdf / weight / color / rem and get_volume() hand out dense views (materialised on first use).
DEFAULT_SPARSE = True

# How ray directions are normalised before the triangle test (normalize(), auxiliary/raytracer/Vector3.h:73-89):
#   "sse"  -- on the host by lib().vl_normalize_rays, the reference's own rsqrtps + Newton step: the device sees the
#             reference's unit vectors bit for bit (x86 hosts; the default there);
#   "ieee" -- on the device with IEEE 1 / sqrt (<= 2 ulp from the reference's; portable).
DEFAULT_NORMALIZE = "sse"


def normalize_rays(rays):
  """Host-side normalize() of the reference (Vector3.h:73-89) on a ray set: numpy / torch f32[R,3] -> numpy f32[3R]."""
  if torch.is_tensor(rays):
    rays = rays.detach().cpu().numpy()
  rays = np.ascontiguousarray(rays, np.float32).reshape(-1)
  out = np.empty_like(rays)
  check(lib().vl_normalize_rays(ctypes.c_void_p(rays.ctypes.data), rays.size // 3, ctypes.c_void_p(out.ctypes.data)))
  return out


def _prepare_rays(rays, normalize, device):
  """-> (flat CUDA tensor, flags for the library)."""
  mode = DEFAULT_NORMALIZE if normalize is None else normalize
  if mode == "sse":
    return _dev(normalize_rays(rays), torch.float32, device).reshape(-1), RAYS_NORMALIZED
  if mode == "ieee":
    return _dev(rays, torch.float32, device).reshape(-1), 0
  if mode == "given":   # the caller's own unit vectors, used as they are
    return _dev(rays, torch.float32, device).reshape(-1), RAYS_NORMALIZED
  raise ValueError("normalize must be 'sse', 'ieee' or 'given'")


def _trace_outputs(n_rays, dev, out, want_ids, alloc=torch.zeros):
  if out is None:
    out = {}
  spec = (("endpoints", 3 * n_rays, torch.float32), ("endcolors", 3 * n_rays, torch.int32),
          ("range", n_rays, torch.float32), ("endrem", n_rays, torch.float32))
  for name, n, dt in spec:
    if name not in out:
      out[name] = alloc(n, dtype=dt, device=dev)  # misses stay 0 (fusion_lidar.py:440-447)
    t = out[name]
    if t.dtype != dt or t.numel() != n or not t.is_contiguous() or t.device != dev:
      raise ValueError("output %s must be a contiguous %s tensor of %d elements on %s" % (name, dt, n, dev))
  if want_ids and "tri_id" not in out:
    out["tri_id"] = torch.empty(n_rays, dtype=torch.int32, device=dev)
  return out


TRACE_PERSISTENT = 16


def trace(bvh, rays, origin, height, out=None, want_ids=True, zero_misses=False, normalize=None, persistent=False):
  """(ii) closest-hit ray cast; the device-resident equivalent of C_Trace
  (auxiliary/raytracer/RayTracerCython.pyx:15-33 -> RayTracer.cpp:56-92).

  rays f32[R,3] (any length; normalised as `normalize` says, see DEFAULT_NORMALIZE), origin f32[3].  Returns dict of flat CUDA
  tensors: endpoints[3R], endcolors[3R], range[R], endrem[R] (written for hits only -- pass
  `out` to keep previous content, or zero_misses=True to have the kernel write 0 for misses)
  and tri_id[R] (original face index, -1 = miss).  persistent=True: the persistent-warp kernel (rays pulled from a
  counter, warp-wide compaction, top of the tree staged in shared memory by TMA; VL_TRACE_PERSISTENT) -- same bits."""
  dev = bvh.blob.device
  rays, ray_flags = _prepare_rays(rays, normalize, dev)
  origin = _dev(origin, torch.float32, dev).reshape(-1)
  n_rays = rays.numel() // 3
  out = _trace_outputs(n_rays, dev, out, want_ids, torch.empty if zero_misses else torch.zeros)
  with torch.cuda.device(dev):
    check(lib().vl_trace(_ptr(bvh.blob), bvh.n_faces, _ptr(rays), _ptr(origin), n_rays, int(height),
                         _ptr(out["endpoints"]), _ptr(out["endcolors"]), _ptr(out["range"]), _ptr(out["endrem"]),
                         _ptr(out.get("tri_id")),
                         (TRACE_ZERO_MISSES if zero_misses else 0) | ray_flags | (TRACE_PERSISTENT if persistent else 0), _stream()))
  return out


class Beams:
  """Direction-space index of one ray set (the target sensor's beams): built once, valid for every scan.

  rays f32[R,3] as MultiSemLaserScan.create_rays returns them (auxiliary/laserscan.py:1092-1119) or any
  other ray set -- the ctrace ABI gives all rays one origin (RayTracer.cpp:116-124), which is all the
  index relies on.  `height` as in C_Trace (width = n_rays // height rays per row are cast).  `normalize`: see
  DEFAULT_NORMALIZE (the directions are normalised once, here)."""

  def __init__(self, rays, height, device=None, normalize=None):
    require_cuda()
    self.rays, ray_flags = _prepare_rays(rays, normalize, device)
    dev = self.rays.device
    self.n_rays = self.rays.numel() // 3
    self.height = int(height)
    self.blob = torch.empty(lib().vl_beams_bytes(self.n_rays, self.height), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
      check(lib().vl_beams_build(_ptr(self.rays), self.n_rays, self.height, _ptr(self.blob), self.blob.numel(), ray_flags,
                                 _stream()))

  def workspace(self, max_faces):
    """Scratch for cast() on meshes of up to max_faces triangles (8 B per ray + 12 B per face)."""
    return torch.empty(lib().vl_cast_workspace_bytes(self.n_rays, int(max_faces)), dtype=torch.uint8, device=self.blob.device)


def cast(beams, verts, faces, colors, rem, origin, out=None, want_ids=True, zero_misses=False, workspace=None,
         check_mesh=True):
  """(ii-b) closest-hit cast of an indexed mesh against indexed beams -- same outputs as
  Bvh(...) + trace(...), i.e. as C_Trace (RayTracerCython.pyx:15-33 -> RayTracer.cpp:19-92), without a
  per-scan tree: the triangles are streamed once through the beam index.  Mesh arrays as for Bvh; colors may also be
  a uint8 CUDA tensor (TsdfDevice.extract_mesh's), read as it is (VL_COLORS_U8).
  check_mesh=True (default) synchronises, raises VlidarError(VL_EBADMESH) on out-of-range face indices or
  VlidarError(VL_ENOSPACE) when the mesh needs more work units than the workspace holds (nothing was written then:
  use Bvh + trace), and adds n_bad_faces / n_active (triangles that can be hit at all) / n_units to the result.
  With check_mesh=False the call is asynchronous and the caller checks lib().vl_cast_status itself."""
  dev = beams.blob.device
  verts = _dev(verts, torch.float32, dev).reshape(-1)
  faces = _dev(faces, torch.int32, dev).reshape(-1) if faces is not None else None   # None: a triangle soup, face f = vertices 3f .. 3f+2
  colors_u8 = torch.is_tensor(colors) and colors.dtype == torch.uint8   # the mesh extraction's colours, used as they are
  colors = _dev(colors, torch.uint8 if colors_u8 else torch.int32, dev).reshape(-1)
  rem = _dev(rem, torch.float32, dev).reshape(-1)
  origin = _dev(origin, torch.float32, dev).reshape(-1)
  n_verts = verts.numel() // 3
  n_faces = faces.numel() // 3 if faces is not None else n_verts // 3
  if colors.numel() != 3 * n_verts or rem.numel() != n_verts:
    raise ValueError("colors must hold 3 ints and rem 1 float per vertex")
  out = _trace_outputs(beams.n_rays, dev, out, want_ids, torch.empty if zero_misses else torch.zeros)
  if workspace is None or workspace.numel() < lib().vl_cast_workspace_bytes(beams.n_rays, n_faces):
    workspace = beams.workspace(n_faces)
  with torch.cuda.device(dev):
    check(lib().vl_cast(_ptr(beams.blob), _ptr(verts), _ptr(faces), _ptr(colors), _ptr(rem), n_verts, n_faces,
                        _ptr(origin), beams.n_rays, beams.height, _ptr(out["endpoints"]), _ptr(out["endcolors"]),
                        _ptr(out["range"]), _ptr(out["endrem"]), _ptr(out.get("tri_id")),
                        (TRACE_ZERO_MISSES if zero_misses else 0) | (COLORS_U8 if colors_u8 else 0), _ptr(workspace),
                        workspace.numel(), _stream()))
    if check_mesh:
      info = (ctypes.c_int * 8)()
      rc = lib().vl_cast_status(_ptr(workspace), _stream(), info)
      out["n_bad_faces"], out["n_active"], out["n_units"] = info[0], info[1], info[2] + (info[3] << 31)
      check(rc)
  return out


def trace_bruteforce(verts, faces, colors, rem, rays, origin, height, out=None, normalize=None):
  """Test aid: same outputs without a BVH (every triangle per ray)."""
  require_cuda()
  verts = _dev(verts, torch.float32).reshape(-1)
  dev = verts.device
  faces = _dev(faces, torch.int32, dev).reshape(-1)
  colors = _dev(colors, torch.int32, dev).reshape(-1)
  rem = _dev(rem, torch.float32, dev).reshape(-1)
  rays, ray_flags = _prepare_rays(rays, normalize, dev)
  origin = _dev(origin, torch.float32, dev).reshape(-1)
  n_rays = rays.numel() // 3
  out = _trace_outputs(n_rays, dev, out, True)
  with torch.cuda.device(dev):
    check(lib().vl_trace_bruteforce(_ptr(verts), _ptr(faces), _ptr(colors), _ptr(rem), verts.numel() // 3,
                                    faces.numel() // 3, _ptr(rays), _ptr(origin), n_rays, int(height),
                                    _ptr(out["endpoints"]), _ptr(out["endcolors"]), _ptr(out["range"]),
                                    _ptr(out["endrem"]), _ptr(out["tri_id"]), ray_flags, _stream()))
  return out


PROJECT_METHODS = {"depth": 0, "pdist": 1, "depthfast": 2}   # VL_PROJECT_* of include/vlidar.h
TSDF_FRESH, TSDF_TABLES_VALID = 1, 2                          # VL_TSDF_* of include/vlidar.h


def project(points, remissions, labels, fov_up, fov_down, H, W, remove=True, workspace=None, beam_angles=None,
            want_bounds=False, out=None, want_keep=True, method="depth"):
  """(iii) spherical range-image projection, the device equivalent of
  LaserScan.do_range_projection_new('depth') + do_label_projection_new
  (auxiliary/laserscan.py:294-391, 672-676).

  points f64[N,3], remissions f32[N], labels u32/i32[N].  Returns dict of CUDA tensors:
  range_image f32[H,W] (0 empty), index i32[H,W] (-1 empty, into the kept points),
  proj_label i32[H,W], proj_remissions f32[H,W] (-1 empty), keep bool[N], n_kept i32[1].
  beam_angles (non-empty sequence): the pitch snapping of laserscan.py:321-327 (vl_project_snap).
  want_bounds: also `bounds` f64[6] = min xyz, max xyz of the kept points (SemLaserScan.get_bnds on the device,
  vl_points_bounds); `bounds` and `n_kept` then share one 64-byte buffer `meta` (a single small D2H for both).
  method: 'depth' (default, :369-391), 'pdist' (:392-416: range_image / index / proj_label of the point nearest to the
  pixel centre) or 'depthfast' (:418-437: range_image starts at -1) -- vl_project_select.
  out: the dict of an earlier call with the same H, W -- its tensors are written again instead of allocating new ones
  (pipeline.ScanPipeline: no allocation per scan); want_keep=False leaves `keep` as the raw uint8 buffer."""
  require_cuda()
  points = _dev(points, torch.float64).reshape(-1)
  dev = points.device
  ba = None
  if beam_angles is not None and len(beam_angles):
    ba_host = np.asarray(beam_angles, np.float64).reshape(-1)
    if not np.isfinite(ba_host).all():   # numpy's argmin would pick the first NaN; refuse instead of guessing
      raise ValueError("beam_angles must be finite")
    ba = _dev(ba_host, torch.float64, dev)
  n = points.numel() // 3
  remissions = _dev(remissions, torch.float32, dev).reshape(-1)
  if torch.is_tensor(labels) and labels.dtype in (torch.int32, getattr(torch, "uint32", torch.int32)):
    labels = labels.to(dev).contiguous().view(torch.int32)
  else:
    labels = _dev(np.asarray(labels.cpu() if torch.is_tensor(labels) else labels).astype(np.uint32).view(np.int32),
                  torch.int32, dev)
  labels = labels.reshape(-1)
  if remissions.numel() != n or labels.numel() != n:
    raise ValueError("points, remissions and labels disagree in length")
  need = lib().vl_project_workspace_bytes(n, H, W)
  if workspace is None or workspace.numel() < need:
    workspace = torch.empty(need, dtype=torch.uint8, device=dev)
  if out is not None and tuple(out["range_image"].shape) == (H, W) and out["range_image"].device == dev and \
      out["_keep_buf"].numel() >= max(n, 1):
    out = dict(out)
    out["keep"] = out["_keep_buf"]
    out["meta"].zero_()
  else:
    out = dict(range_image=torch.empty((H, W), dtype=torch.float32, device=dev),
               index=torch.empty((H, W), dtype=torch.int32, device=dev),
               proj_label=torch.empty((H, W), dtype=torch.int32, device=dev),
               proj_remissions=torch.empty((H, W), dtype=torch.float32, device=dev),
               keep=torch.empty(max(n, 1) + max(n, 1) // 4, dtype=torch.uint8, device=dev))
    out["_keep_buf"] = out["keep"]
    meta = torch.zeros(64, dtype=torch.uint8, device=dev)
    out["meta"], out["bounds"], out["n_kept"] = meta, meta[:48].view(torch.float64), meta[48:52].view(torch.int32)
  with torch.cuda.device(dev):
    check(lib().vl_project_select(_ptr(points), _ptr(remissions), _ptr(labels), n, float(fov_up), float(fov_down), H, W,
                                1 if remove else 0, _ptr(ba) if ba is not None else None,
                                ba.numel() if ba is not None else 0, PROJECT_METHODS[method], _ptr(out["range_image"]), _ptr(out["index"]),
                                _ptr(out["proj_label"]), _ptr(out["proj_remissions"]), _ptr(out["keep"]),
                                _ptr(out["n_kept"]), _ptr(workspace), workspace.numel(), _stream()))
    if want_bounds:
      check(lib().vl_points_bounds(_ptr(points), _ptr(out["keep"]), n, _ptr(out["bounds"]), _stream()))
  out["keep"] = out["keep"][:n].bool() if want_keep else out["keep"][:n]
  out["workspace"] = workspace
  return out


def reverse_project(depth_im, proj_x, proj_y, fov_up, fov_down):
  """Pixel coordinates + depth -> xyz on the device, the `cp` adaption's back projection
  (LaserScan.do_reverse_projection_new, auxiliary/laserscan.py:475-501).  depth_im f32[H,W], proj_x / proj_y [H,W]
  (float image coordinates or the clamped integer ones).  Returns back_points f64[H*W,3] (CUDA tensor)."""
  require_cuda()
  depth_im = _dev(depth_im, torch.float32)
  dev = depth_im.device
  H, W = depth_im.shape
  px, py = _dev(proj_x, torch.float64, dev), _dev(proj_y, torch.float64, dev)
  if px.numel() != H * W or py.numel() != H * W:
    raise ValueError("proj_x / proj_y must have one entry per pixel")
  out = torch.empty((H * W, 3), dtype=torch.float64, device=dev)
  with torch.cuda.device(dev):
    check(lib().vl_reverse_project(_ptr(depth_im), _ptr(px), _ptr(py), int(H), int(W), float(fov_up), float(fov_down),
                                   _ptr(out), _stream()))
  return out


class TsdfDevice:
  """(iv) the four TSDF volumes resident in HBM + the integrate kernel.

  Device-side core of auxiliary.fusion_lidar.TSDFVolume (auxiliary/fusion_lidar.py:23-63,
  252-287); the reference-shaped class lives in lidar_transfer_b200/auxiliary/fusion_lidar.py."""

  def __init__(self, vol_dim, vol_origin, voxel_size, fov_up, fov_down, device=None, sparse=None):
    require_cuda()
    self.sparse = DEFAULT_SPARSE if sparse is None else bool(sparse)
    self.dim = tuple(int(v) for v in vol_dim)
    self.origin = np.asarray(vol_origin, np.float32).copy()
    self.voxel_size = float(np.float32(voxel_size))
    self.trunc_margin = float(np.float32(voxel_size * 5))  # fusion_lidar.py:31, cast at :277-280
    self.fov_up = float(np.float32(fov_up))
    self.fov_down = float(np.float32(fov_down))
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = self.dim[0] * self.dim[1] * self.dim[2]
    if n >= 2 ** 31:
      raise ValueError("volume of %d voxels exceeds the reference kernel's int voxel index" % n)
    self._store = self._acquire(n, dev)
    self._vols = tuple(f[:n].view(self.dim) for f in self._store["flat"])
    if self.dim[2] > 32767:
      self.sparse = False
    self.reset()

  # The reference allocates (and uploads) four new volumes per TSDFVolume, i.e. per scan (laserscan.py:968,
  # fusion_lidar.py:48-63).  Here the storage of a volume that has been dropped is kept for the next one of a similar
  # size (bounds are clipped to each scan's points, so sizes differ slightly from scan to scan): no cudaMalloc / cudaFree
  # -- and none of their device synchronisations -- in a batch.  At most _POOL_KEEP idle stores are kept.
  _pool = []
  _POOL_KEEP = 2

  @classmethod
  def _acquire(cls, n, dev):
    best = None
    for st in cls._pool:
      if st["dev"] == dev and n <= st["cap"] <= max(2 * n, n + (1 << 24)) and (best is None or st["cap"] < best["cap"]):
        best = st
    if best is not None:
      cls._pool.remove(best)
      return best
    cap = n + n // 16 + 1024    # a little head room for the next scan's slightly larger box
    return dict(dev=dev, cap=cap, flat=tuple(torch.empty(cap, dtype=torch.float32, device=dev) for _ in range(4)), ws={})

  def release(self):
    """Hands the storage back for reuse; the volumes must not be used afterwards."""
    st, self._store = getattr(self, "_store", None), None
    if st is not None:
      self._vols = None
      TsdfDevice._pool.append(st)
      while len(TsdfDevice._pool) > TsdfDevice._POOL_KEEP:
        TsdfDevice._pool.pop(0)

  def __del__(self):
    try:
      self.release()
    except Exception:
      pass

  def _workspace(self, name, need, dev):
    ws = self._store["ws"].get(name)
    if ws is None or ws.numel() < need:
      ws = self._store["ws"][name] = torch.empty(need, dtype=torch.uint8, device=dev)
    return ws

  def reset(self):
    """Back to the initial state (tsdf 1, weight / colour / remission 0, fusion_lidar.py:48-63).  Lazy: the first
    integrate() writes the initial values of the voxels it does not update in the same pass
    (vl_tsdf_init_integrate); anything else that looks at the volumes first materialises them (vl_tsdf_init)."""
    self._fresh = True
    self._dense = False   # sparse volume: every voxel materialised (hull = whole columns)?

  def _hull(self):
    """Per z column the interval of voxels that exist (vl_tsdf_sparse_integrate), i32[dx * dy] on the device."""
    n_cols = self.dim[0] * self.dim[1]
    return self._workspace("hull", 4 * n_cols, self._vols[0].device)[:4 * n_cols].view(torch.int32)

  def _materialise(self):
    """Dense volumes: what the properties below, get_volume() and the dense API see."""
    t, w, c, r = self._vols
    if self._fresh:
      self._fresh = False
      with torch.cuda.device(t.device):
        check(lib().vl_tsdf_init(_ptr(t), _ptr(w), _ptr(c), _ptr(r), t.numel(), _stream()))
      if self.sparse:
        self._hull().fill_((self.dim[2] - 1) << 16)
        self._dense = True
    elif self.sparse and not self._dense:
      with torch.cuda.device(t.device):
        check(lib().vl_tsdf_densify(_ptr(t), _ptr(w), _ptr(c), _ptr(r), self.dim[0], self.dim[1], self.dim[2],
                                    _ptr(self._hull()), _stream()))
      self._dense = True

  tsdf = property(lambda self: (self._materialise(), self._vols[0])[1])
  weight = property(lambda self: (self._materialise(), self._vols[1])[1])
  color = property(lambda self: (self._materialise(), self._vols[2])[1])
  rem = property(lambda self: (self._materialise(), self._vols[3])[1])

  def integrate(self, color_im, depth_im, rem_im, obs_weight=1.0, use_column_table=True):
    """color_im: folded single-channel image (label * 65536), depth_im, rem_im: f32[H,W].
    use_column_table: vl_tsdf_integrate_ws (per-column pixel table, same bits) instead of vl_tsdf_integrate."""
    dev = self._vols[0].device
    color_im = _dev(color_im, torch.float32, dev)
    depth_im = _dev(depth_im, torch.float32, dev)
    rem_im = _dev(rem_im, torch.float32, dev)
    im_h, im_w = depth_im.shape
    origin = (ctypes.c_float * 3)(*[float(v) for v in self.origin])
    if self.sparse and use_column_table:
      # sparse volume: no reset pass, only the voxels inside the columns' hulls exist (vl_tsdf_sparse_integrate)
      need = lib().vl_tsdf_sparse_workspace_bytes(self.dim[0], self.dim[1], int(im_h), int(im_w))
      ws = self._workspace("columns", need, dev)
      # the geometry tables in the workspace (pixel column per z column, tangents per image row) depend on the volume
      # geometry, the field of view and the image size only: rebuilt when any of those -- or the workspace -- changes
      key = (ws.data_ptr(), self.dim, self.origin.tobytes(), self.voxel_size, self.fov_up, self.fov_down, int(im_h), int(im_w),
             torch.cuda.current_stream(dev).cuda_stream)
      flags = (TSDF_FRESH if self._fresh else 0) | (TSDF_TABLES_VALID if self._store.get("tables_key") == key else 0)
      self._store["tables_key"] = None
      with torch.cuda.device(dev):
        check(lib().vl_tsdf_sparse_integrate(_ptr(self._vols[0]), _ptr(self._vols[1]), _ptr(self._vols[2]), _ptr(self._vols[3]),
                                             self.dim[0], self.dim[1], self.dim[2], origin, self.voxel_size, self.trunc_margin,
                                             float(np.float32(obs_weight)), self.fov_up, self.fov_down, _ptr(color_im),
                                             _ptr(depth_im), _ptr(rem_im), int(im_h), int(im_w), _ptr(self._hull()),
                                             flags, _ptr(ws), ws.numel(), _stream()))
      self._store["tables_key"] = key
      if self._fresh:
        self._fresh, self._dense = False, False
      return
    self._store["tables_key"] = None   # the dense calls lay the workspace out their own way
    fused = self._fresh and use_column_table   # first integration into a fresh volume: one pass
    if not fused:
      self._materialise()
    if use_column_table:
      need = lib().vl_tsdf_fresh_workspace_bytes(self.dim[0], self.dim[1], int(im_h), int(im_w))  # column table + shell image
      ws = self._workspace("columns", need, dev)
    with torch.cuda.device(dev):
      if use_column_table:
        fn = lib().vl_tsdf_init_integrate if fused else lib().vl_tsdf_integrate_ws
        check(fn(_ptr(self._vols[0]), _ptr(self._vols[1]), _ptr(self._vols[2]), _ptr(self._vols[3]),
                 self.dim[0], self.dim[1], self.dim[2], origin, self.voxel_size,
                 self.trunc_margin, float(np.float32(obs_weight)), self.fov_up, self.fov_down,
                 _ptr(color_im), _ptr(depth_im), _ptr(rem_im), int(im_h), int(im_w), _ptr(ws), ws.numel(), _stream()))
        if fused:
          self._fresh = False   # only once the fused reset + integration has been accepted: a failed call leaves the volume "fresh"
      else:
        check(lib().vl_tsdf_integrate(_ptr(self.tsdf), _ptr(self.weight), _ptr(self.color), _ptr(self.rem),
                                      self.dim[0], self.dim[1], self.dim[2], origin, self.voxel_size,
                                      self.trunc_margin, float(np.float32(obs_weight)), self.fov_up, self.fov_down,
                                      _ptr(color_im), _ptr(depth_im), _ptr(rem_im), int(im_h), int(im_w), _stream()))


  def extract_mesh(self, level=0.0, want_norms=True, want_faces=True):
    """(v) iso-surface of the TSDF volume at `level` + per-vertex colour / remission lookup on the device;
    the device-resident equivalent of TSDFVolume.get_mesh (auxiliary/fusion_lidar.py:403-424).
    Returns dict(verts f32[N_v,3] world frame, faces i32[N_t,3], norms f32[N_v,3], colors u8[N_v,3],
    rem f32[N_v]) -- a triangle soup, N_v = 3 N_t.  Synchronises once (the triangle count).
    want_faces=False: faces is None (the index array of a soup is 0 .. 3 N_t - 1; cast() takes the soup without it)."""
    return self.extract_mesh_finish(self.extract_mesh_begin(level), want_norms, want_faces=want_faces)

  def extract_mesh_begin(self, level=0.0):
    """First half of extract_mesh: the counting sweep is enqueued on the current stream and the two totals start their
    way to pinned host memory; nothing waits.  pipeline.ScanPipeline puts another scan's kernels between the halves."""
    hull = None
    if self.sparse and not self._fresh and not self._dense and level <= 1.0 and self.dim[2] + 64 <= 2048:
      hull = self._hull()    # sparse volume: the mesh kernels read inside the hulls only
      vols = self._vols
    else:
      vols = (self.tsdf, self.weight, self.color, self.rem)
    dev = vols[0].device
    need = lib().vl_mesh_workspace_bytes(self.dim[0], self.dim[1], self.dim[2])
    ws = self._workspace("mesh", need, dev)
    totals = self._workspace("mesh_totals", 16, dev)[:16].view(torch.int64)
    totals.zero_()
    with torch.cuda.device(dev):
      if hull is not None:
        rc = lib().vl_mesh_count_sparse(_ptr(vols[0]), self.dim[0], self.dim[1], self.dim[2], float(level), _ptr(hull), _ptr(ws),
                                        ws.numel(), _ptr(totals), _stream())
        if rc == _lib.VL_EINVAL:   # e.g. a debug sweep that needs the dense volume (vl_debug_mesh_scalar(2 / 3))
          hull, vols = None, (self.tsdf, self.weight, self.color, self.rem)
        else:
          check(rc)
      if hull is None:
        check(lib().vl_mesh_count(_ptr(vols[0]), self.dim[0], self.dim[1], self.dim[2], float(level), _ptr(ws),
                                  ws.numel(), _ptr(totals), _stream()))
    if getattr(self, "_h_totals", None) is None:
      self._h_totals = torch.zeros(2, dtype=torch.int64).pin_memory()
      self._ev_totals = torch.cuda.Event()
    self._h_totals.copy_(totals, non_blocking=True)
    self._ev_totals.record(torch.cuda.current_stream(dev))
    return dict(level=float(level), hull=hull, vols=vols, ws=ws, dev=dev)

  def extract_mesh_finish(self, ctx, want_norms=True, buffers=None, want_faces=True):
    """Second half: waits for the totals (the one host synchronisation of the extraction), allocates the mesh and
    enqueues the emit on the current stream (the stream extract_mesh_begin ran on).  buffers: a dict the caller keeps
    between calls -- the mesh arrays are then views into grow-only storage held there (valid until the next call with the
    same dict) instead of fresh allocations."""
    hull, vols, ws, dev, level = ctx["hull"], ctx["vols"], ctx["ws"], ctx["dev"], ctx["level"]
    self._ev_totals.synchronize()
    n_t, n_a = (int(v) for v in self._h_totals.tolist())
    origin = (ctypes.c_float * 3)(*[float(v) for v in self.origin])
    with torch.cuda.device(dev):
      list_bytes = lib().vl_mesh_list_bytes(n_t, n_a)
      if buffers is not None:
        if buffers.get("cap_t", -1) < n_t or (want_norms and buffers.get("norms") is None) or (want_faces and buffers.get("faces") is None):
          cap = n_t + n_t // 4 + 1024
          buffers.update(cap_t=cap, verts=torch.empty((3 * cap, 3), dtype=torch.float32, device=dev),
                         faces=torch.empty((cap, 3), dtype=torch.int32, device=dev) if want_faces else None,
                         norms=torch.empty((3 * cap, 3), dtype=torch.float32, device=dev) if want_norms else None,
                         colors=torch.empty((3 * cap, 3), dtype=torch.uint8, device=dev),
                         rem=torch.empty(3 * cap, dtype=torch.float32, device=dev))
        if buffers.get("active") is None or buffers["active"].numel() < list_bytes:
          buffers["active"] = torch.empty(list_bytes + list_bytes // 4, dtype=torch.uint8, device=dev)
        out = dict(verts=buffers["verts"][:3 * n_t], faces=buffers["faces"][:n_t] if want_faces else None,
                   norms=buffers["norms"][:3 * n_t] if want_norms else None, colors=buffers["colors"][:3 * n_t],
                   rem=buffers["rem"][:3 * n_t])
        active = buffers["active"]
      else:
        out = dict(verts=torch.empty((3 * n_t, 3), dtype=torch.float32, device=dev),
                   faces=torch.empty((n_t, 3), dtype=torch.int32, device=dev) if want_faces else None,
                   norms=torch.empty((3 * n_t, 3), dtype=torch.float32, device=dev) if want_norms else None,
                   colors=torch.empty((3 * n_t, 3), dtype=torch.uint8, device=dev),
                   rem=torch.empty(3 * n_t, dtype=torch.float32, device=dev))
        active = torch.empty(list_bytes, dtype=torch.uint8, device=dev)   # scratch of vl_mesh_emit
      if hull is not None:
        check(lib().vl_mesh_emit_sparse(_ptr(vols[0]), _ptr(vols[2]), _ptr(vols[3]), self.dim[0], self.dim[1], self.dim[2],
                                        float(level), self.voxel_size, origin, _ptr(hull), _ptr(ws), ws.numel(), n_t, n_a,
                                        _ptr(active), _ptr(out["verts"]), _ptr(out["faces"]), _ptr(out["norms"]),
                                        _ptr(out["colors"]), _ptr(out["rem"]), _stream()))
      else:
        check(lib().vl_mesh_emit(_ptr(vols[0]), _ptr(vols[2]), _ptr(vols[3]), self.dim[0], self.dim[1], self.dim[2],
                                 float(level), self.voxel_size, origin, _ptr(ws), ws.numel(), n_t, n_a, _ptr(active),
                                 _ptr(out["verts"]), _ptr(out["faces"]), _ptr(out["norms"]), _ptr(out["colors"]),
                                 _ptr(out["rem"]), _stream()))
      out["n_active_cubes"], out["n_tris"] = n_a, n_t
      out["_scratch"] = active   # read by the emit kernel in flight: lives as long as the mesh does
    return out


def compare(source_color, target_color, source_label, target_label, source_range, target_range, source_rem, target_rem,
            nclasses, workspace=None):
  """(vi) identity re-render metrics on the device: the masks, difference images, label renumbering and confusion
  matrix of compare() (auxiliary/laserscan.py:1181-1301) + iouEval.addBatch (auxiliary/np_ioueval.py:32-45); IoU and
  accuracy (np_ioueval.py:47-70) are evaluated on the nclasses x nclasses matrix on the host.

  Images as numpy arrays or torch tensors (any device): colours [H,W,3], labels [H,W], ranges / remissions [H,W].
  Returns dict(label_diff f32[H,W,3], range_diff f32[H,W], rem_diff f32[H,W] (CUDA tensors), conf i64[n,n] (numpy),
  n_present, m_iou, m_acc, mse).  Synchronises once (the confusion matrix)."""
  require_cuda()
  sc = _dev(source_color, torch.float32)
  dev = sc.device
  shape = tuple(sc.shape[:-1])
  n = int(np.prod(shape))
  tc = _dev(target_color, torch.float32, dev)
  sl, tl = _dev(source_label, torch.int32, dev), _dev(target_label, torch.int32, dev)
  sr, tr = _dev(source_range, torch.float32, dev), _dev(target_range, torch.float32, dev)
  sm, tm = _dev(source_rem, torch.float32, dev), _dev(target_rem, torch.float32, dev)
  for name, t, k in (("target_color", tc, 3 * n), ("source_label", sl, n), ("target_label", tl, n), ("source_range", sr, n),
                     ("target_range", tr, n), ("source_rem", sm, n), ("target_rem", tm, n)):
    if t.numel() != k:
      raise ValueError("%s has %d elements, expected %d" % (name, t.numel(), k))
  nclasses = int(nclasses)
  need = lib().vl_compare_workspace_bytes(n)
  if workspace is None or workspace.numel() < need:
    workspace = torch.empty(need, dtype=torch.uint8, device=dev)
  out = dict(label_diff=torch.empty(shape + (3,), dtype=torch.float32, device=dev),
             range_diff=torch.empty(shape, dtype=torch.float32, device=dev),
             rem_diff=torch.empty(shape, dtype=torch.float32, device=dev))
  conf = torch.empty((nclasses, nclasses), dtype=torch.int64, device=dev)
  with torch.cuda.device(dev):
    check(lib().vl_compare(_ptr(sc), _ptr(tc), _ptr(sl), _ptr(tl), _ptr(sr), _ptr(tr), _ptr(sm), _ptr(tm), n, nclasses,
                           _ptr(out["label_diff"]), _ptr(out["range_diff"]), _ptr(out["rem_diff"]), _ptr(conf),
                           _ptr(workspace), workspace.numel(), _stream()))
    info = (ctypes.c_int * 8)()
    sq = ctypes.c_double(0.0)
    check(lib().vl_compare_status(_ptr(workspace), _stream(), info, ctypes.byref(sq)))
  conf = conf.cpu().numpy()
  k = info[0]
  # np_ioueval.py:12-17, 47-70 with ignore = the class indices that no renumbered label uses (laserscan.py:1225-1230)
  ignore = np.arange(k, nclasses)
  include = np.arange(0, k)
  c = conf.copy()
  c[ignore] = 0
  c[:, ignore] = 0
  tp = np.diag(c)
  fp, fn = c.sum(axis=1) - tp, c.sum(axis=0) - tp
  union = tp + fp + fn + 1e-15
  out.update(conf=conf, n_present=k, m_iou=(tp[include] / union[include]).mean(),
             m_acc=tp.sum() / (tp[include].sum() + fp[include].sum() + 1e-15), mse=sq.value / max(n, 1),
             workspace=workspace)
  return out


def ctrace_host(rays, origin, verts, faces, colors, rem, height, outputs=None, want_ids=False, method=None,
                normalize=None):
  """The reference-compatible HOST-pointer entry point (extern "C" ctrace / vl_ctrace_ids) on numpy
  buffers: H2D, build, trace, D2H inside the call.  outputs: dict of preallocated numpy arrays
  (endpoints, endcolors, range, endrem) updated in place for hits."""
  require_cuda()
  f32, i32 = np.float32, np.int32

  def chk(a, dt, name):
    if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"]):
      raise ValueError("%s must be a C-contiguous numpy array of %s" % (name, np.dtype(dt).name))
    return a

  rays = chk(rays, f32, "rays").reshape(-1)
  origin = chk(origin, f32, "origin").reshape(-1)
  verts = chk(verts, f32, "verts").reshape(-1)
  faces = chk(faces, i32, "faces").reshape(-1)
  colors = chk(colors, i32, "colors").reshape(-1)
  rem = chk(rem, f32, "rem").reshape(-1)
  n_rays = rays.size // 3
  if outputs is None:
    outputs = dict(endpoints=np.zeros(3 * n_rays, f32), endcolors=np.zeros(3 * n_rays, i32),
                   range=np.zeros(n_rays, f32), endrem=np.zeros(n_rays, f32))
  tri_id = np.empty(n_rays, i32) if want_ids else None
  p = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)
  if method is not None:  # "cast" (default of the library) or "lbvh"
    lib().vl_ctrace_method({"cast": 0, "lbvh": 1}[method])
  lib().vl_ctrace_normalize({"sse": 0, "ieee": 1}[DEFAULT_NORMALIZE if normalize is None else normalize])
  check(lib().vl_ctrace_ids(p(rays), p(origin), p(verts), p(faces), p(colors), p(rem), n_rays, verts.size // 3,
                            faces.size // 3, int(height), p(chk(outputs["endpoints"], f32, "endpoints")),
                            p(chk(outputs["endcolors"], i32, "endcolors")), p(chk(outputs["range"], f32, "range")),
                            p(chk(outputs["endrem"], f32, "endrem")), p(tri_id)))
  if want_ids:
    outputs["tri_id"] = tri_id
  return outputs
