"""Python front-end of the CHECKERS.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (lidar_transfer_b200) never does.

Two families of checkers live here:

* ``oracle.*``   -- oracle/liboracle.so, our C restatement (oracle/vl_oracle.c) plus the
                    numpy restatements below.  Builds anywhere gcc exists.
* ``oracle.ref_*`` -- oracle/_ref/*.so, the reference itself compiled from its own
                    sources (oracle/Makefile `ref` target; only buildable where
                    /root/reference exists, the binaries travel to the GPU box).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

NORMALIZE_SSE = 1  # Vector3.h:73-89 rsqrtps + Newton step (bit-exact vs the reference on x86)
BRUTE_FORCE = 2    # every triangle, tie-break (min t, min triangle index)
MIN_ID_TIES = 4    # reference BVH, exact-t ties to the smaller triangle index

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i64p = ctypes.POINTER(ctypes.c_longlong)


def build(ref=None):
  """Compile liboracle.so and -- when the reference checkout is present -- oracle/_ref."""
  if ref is None:
    ref = os.path.isdir(REFERENCE_ROOT)
  subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
  if ref:
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_libs = {}


def _lib(path):
  if path not in _libs:
    if not os.path.exists(path):
      raise FileNotFoundError(
          "%s missing: run `make -C oracle` (oracle) / `make -C oracle ref` (reference build)" % path)
    _libs[path] = ctypes.CDLL(path)
  return _libs[path]


def have_ref(name="libref_raytracer_nofma.so"):
  return os.path.exists(os.path.join(_REF_DIR, name))


def _p(a, t):
  return a.ctypes.data_as(t)


def _prep_trace(rays, origin, verts, faces, colors, rem):
  rays = np.ascontiguousarray(rays, np.float32).reshape(-1)
  origin = np.ascontiguousarray(origin, np.float32).reshape(-1)
  verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
  faces = np.ascontiguousarray(faces, np.int32).reshape(-1)
  colors = np.ascontiguousarray(colors, np.int32).reshape(-1)
  rem = np.ascontiguousarray(rem, np.float32).reshape(-1)
  n_rays = rays.size // 3
  out = dict(endpoints=np.zeros(n_rays * 3, np.float32), endcolors=np.zeros(n_rays * 3, np.int32),
             range=np.zeros(n_rays, np.float32), endrem=np.zeros(n_rays, np.float32),
             tri_id=np.full(n_rays, -1, np.int32))
  return rays, origin, verts, faces, colors, rem, n_rays, out


def trace(rays, origin, verts, faces, colors, rem, height, flags=0):
  """vlo_trace: restatement of trace()/ctrace, auxiliary/raytracer/RayTracer.cpp:19-124."""
  lib = _lib(os.path.join(_HERE, "liboracle.so"))
  rays, origin, verts, faces, colors, rem, n_rays, out = _prep_trace(rays, origin, verts, faces, colors, rem)
  stats = np.zeros(4, np.int32)
  lib.vlo_trace(_p(rays, _f32p), _p(origin, _f32p), _p(verts, _f32p), _p(faces, _i32p), _p(colors, _i32p),
                _p(rem, _f32p), ctypes.c_int(n_rays), ctypes.c_int(verts.size // 3),
                ctypes.c_int(faces.size // 3), ctypes.c_int(height),
                _p(out["endpoints"], _f32p), _p(out["endcolors"], _i32p), _p(out["range"], _f32p),
                _p(out["endrem"], _f32p), _p(out["tri_id"], _i32p), ctypes.c_uint(flags), _p(stats, _i32p))
  out["n_nodes"] = int(stats[0])
  return out


def normalize_rays(rays, flags=NORMALIZE_SSE):
  """vlo_normalize_rays: normalize() of Vector3.h:73-89 alone (SSE mode = the reference's bits)."""
  lib = _lib(os.path.join(_HERE, "liboracle.so"))
  rays = np.ascontiguousarray(rays, np.float32).reshape(-1)
  out = np.empty_like(rays)
  lib.vlo_normalize_rays(_p(rays, _f32p), ctypes.c_long(rays.size // 3), ctypes.c_uint(flags), _p(out, _f32p))
  return out


def ray_triangle(ray, origin, v0, v1, v2, flags=NORMALIZE_SSE):
  """vlo_ray_triangle: t of one ray against one triangle under the reference arithmetic, or None on a miss."""
  lib = _lib(os.path.join(_HERE, "liboracle.so"))
  a = [np.ascontiguousarray(x, np.float32).reshape(3) for x in (ray, origin, v0, v1, v2)]
  t = ctypes.c_float(0.0)
  hit = lib.vlo_ray_triangle(*[_p(x, _f32p) for x in a], ctypes.c_uint(flags), ctypes.byref(t))
  return np.float32(t.value) if hit else None


def ref_ctrace(rays, origin, verts, faces, colors, rem, height, variant="nofma", ids=False):
  """The reference's own extern "C" ctrace (RayTracer.cpp:116-124) compiled from its sources.

  variant "nofma" = canonical parity build (-ffp-contract=off), "fma" = reference flags as shipped.
  ids=True uses the triangle-id harness (oracle/ref_ids_harness.cpp, nofma only).
  """
  rays, origin, verts, faces, colors, rem, n_rays, out = _prep_trace(rays, origin, verts, faces, colors, rem)
  args = [_p(rays, _f32p), _p(origin, _f32p), _p(verts, _f32p), _p(faces, _i32p), _p(colors, _i32p),
          _p(rem, _f32p), ctypes.c_int(n_rays), ctypes.c_int(verts.size // 3), ctypes.c_int(faces.size // 3),
          ctypes.c_int(height), _p(out["endpoints"], _f32p), _p(out["endcolors"], _i32p),
          _p(out["range"], _f32p), _p(out["endrem"], _f32p)]
  if ids:
    lib = _lib(os.path.join(_REF_DIR, "libref_ids_nofma.so"))
    lib.ctrace_ids(*args, _p(out["tri_id"], _i32p))
  else:
    name = "libref_raytracer_nofma.so" if variant == "nofma" else "libref_raytracer.so"
    lib = _lib(os.path.join(_REF_DIR, name))
    lib.ctrace(*args)
    out.pop("tri_id")
  return out


def create_rays(fov_up, fov_down, H, W):
  """Restates MultiSemLaserScan.create_rays, auxiliary/laserscan.py:1092-1119."""
  initial = 180
  yaw_angles = (np.linspace(0, 360, W) + initial)
  larger = yaw_angles > 360
  yaw_angles[larger] -= 360
  yaw = yaw_angles / 180. * np.pi
  pitch = np.linspace(fov_up, fov_down, H) / 180. * np.pi
  pitch = np.pi / 2 - pitch
  beams = np.empty((H, W, 3), np.float64)
  for i, p in enumerate(pitch):
    beams[i, :, 0] = np.sin(p) * np.cos(-yaw)
    beams[i, :, 1] = np.sin(p) * np.sin(-yaw)
    beams[i, :, 2] = np.cos(p) * np.ones(yaw.shape)
  return np.ascontiguousarray(beams.reshape(W * H, -1).astype(np.float32))


def project(points, remissions, labels, fov_up, fov_down, H, W, remove=True, beam_angles=None):
  """vlo_project[_snap]: do_range_projection_new('depth') + do_label_projection_new,
  auxiliary/laserscan.py:294-391, 672-676; beam_angles (a non-empty list) turns on the pitch snapping of
  :321-327."""
  lib = _lib(os.path.join(_HERE, "liboracle.so"))
  lib.vlo_project_snap.restype = ctypes.c_long
  ba = np.ascontiguousarray(beam_angles if beam_angles is not None else [], np.float64).reshape(-1)
  points = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
  n = points.shape[0]
  remissions = np.ascontiguousarray(remissions, np.float32)
  labels = np.ascontiguousarray(labels, np.uint32)
  out = dict(range_image=np.empty(H * W, np.float32), index=np.empty(H * W, np.int32),
             proj_label=np.empty(H * W, np.int32), proj_remissions=np.empty(H * W, np.float32),
             keep=np.empty(n, np.uint8))
  kept = lib.vlo_project_snap(_p(points, _f64p), _p(remissions, _f32p), _p(labels, _u32p), ctypes.c_long(n),
                              ctypes.c_double(fov_up), ctypes.c_double(fov_down), ctypes.c_int(H), ctypes.c_int(W),
                              ctypes.c_int(1 if remove else 0), _p(ba, _f64p), ctypes.c_int(ba.size),
                              _p(out["range_image"], _f32p), _p(out["index"], _i32p),
                              _p(out["proj_label"], _i32p), _p(out["proj_remissions"], _f32p), _p(out["keep"], _u8p))
  for k in ("range_image", "index", "proj_label", "proj_remissions"):
    out[k] = out[k].reshape(H, W)
  out["n_kept"] = int(kept)
  out["keep"] = out["keep"].astype(bool)
  return out


def project_numpy(points, remissions, labels, fov_up, fov_down, H, W, remove=True, beam_angles=None, method="depth"):
  """Pure-numpy/Python-loop restatement of the same projection (small inputs only);
  line-for-line semantics of auxiliary/laserscan.py:294-391 used to cross-check vlo_project.
  method: 'depth' (:369-391), 'pdist' (:392-416: the point nearest to the pixel centre, compared against a float32
  image like 'depth'; proj_remissions is never written) or 'depthfast' (:418-437: descending argsort + fancy-index
  assignment, last write wins = the smallest depth; images start at -1; among EQUAL depths the reference's winner is
  whatever numpy's unstable argsort leaves last -- here the smallest index).  Pinned by tests/golden/golden_methods_v1.npz."""
  points = np.asarray(points, np.float64).reshape(-1, 3)
  remissions = np.asarray(remissions, np.float32)
  labels = np.asarray(labels, np.uint32)
  fu = fov_up / 180.0 * np.pi
  fd = fov_down / 180.0 * np.pi
  fov = abs(fd) + abs(fu)
  depth = np.linalg.norm(points, 2, axis=1)
  k0 = depth != 0
  depth, points, remissions, labels = depth[k0], points[k0], remissions[k0], labels[k0]
  yaw = -np.arctan2(points[:, 1], points[:, 0])
  pitch = np.arcsin(points[:, 2] / depth)
  if beam_angles is not None and len(beam_angles):  # :321-327
    ba = np.asarray(beam_angles, np.float64)
    for i in range(len(pitch)):
      pitch[i] = ba[np.abs(pitch[i] - ba).argmin()]
  proj_x = 0.5 * (yaw / np.pi + 1.0)
  proj_y = 1.0 - (pitch + abs(fd)) / fov
  keep = k0.copy()
  if remove:
    k1 = (proj_y >= 0) & (proj_y <= 1)
    depth, proj_x, proj_y, remissions, labels = depth[k1], proj_x[k1], proj_y[k1], remissions[k1], labels[k1]
    keep[np.flatnonzero(k0)[~k1]] = False
  proj_x = proj_x * W
  proj_y = proj_y * H
  px = np.maximum(0, np.minimum(W - 1, np.floor(proj_x))).astype(np.int32)
  py = np.maximum(0, np.minimum(H - 1, np.floor(proj_y))).astype(np.int32)
  index = np.full((H, W), -1, np.int32)
  range_image = np.zeros((H, W), np.float32)
  proj_rem = np.full((H, W), -1, np.float32)
  if method == "depth":
    for i in range(len(depth)):
      if depth[i] < range_image[py[i], px[i]] or index[py[i], px[i]] == -1:
        range_image[py[i], px[i]] = depth[i]
        index[py[i], px[i]] = i
        proj_rem[py[i], px[i]] = remissions[i]
  elif method == "pdist":
    dist_image = np.full((H, W), 1000, np.float32)
    for i in range(len(depth)):
      dist = np.linalg.norm(np.array([proj_y[i], proj_x[i]]) - np.array([py[i] + 0.5, px[i] + 0.5]))
      if dist < dist_image[py[i], px[i]]:
        dist_image[py[i], px[i]] = dist
        range_image[py[i], px[i]] = depth[i]
        index[py[i], px[i]] = i
  elif method == "depthfast":
    range_image[:] = -1
    order = np.lexsort((np.arange(len(depth)), depth))[::-1]   # depth descending; equal depths: the smaller index is written last
    range_image[py[order], px[order]] = depth[order]
    proj_rem[py[order], px[order]] = remissions[order]
    index[py[order], px[order]] = np.arange(len(depth))[order]
  else:
    raise ValueError(method)
  proj_label = np.zeros((H, W), np.int32)
  mask = index >= 0
  proj_label[mask] = labels[index[mask]]
  return dict(range_image=range_image, index=index, proj_label=proj_label, proj_remissions=proj_rem,
              keep=keep, n_kept=int(len(depth)))


def _tsdf_args(vol_dim, vol_origin, voxel_size, im_h, im_w, trunc, obs_weight, fov_up, fov_down):
  vd = np.asarray(vol_dim, np.float32).copy()
  vo = np.asarray(vol_origin, np.float32).copy()
  other = np.asarray([0, voxel_size, im_h, im_w, trunc, obs_weight, fov_up, fov_down], np.float32)
  return vd, vo, other


def tsdf_new_volume(vol_dim):
  """Fresh volumes as TSDFVolume.__init__ makes them, auxiliary/fusion_lidar.py:48-52."""
  vol_dim = tuple(int(v) for v in vol_dim)
  return dict(tsdf=np.ones(vol_dim, np.float32), weight=np.zeros(vol_dim, np.float32),
              color=np.zeros(vol_dim, np.float32), rem=np.zeros(vol_dim, np.float32))


def label_to_color_im(proj_label):
  """integrate()'s fold of the 3-channel image, fusion_lidar.py:259-264, for
  proj_label3[:, :, 0] = label (laserscan.py:970-971)."""
  lab = np.asarray(proj_label).astype(np.float32)
  return np.floor(lab * 256 * 256 + np.float32(0) * 256 + np.float32(0)).astype(np.float32)


def tsdf_integrate(vol, vol_origin, voxel_size, color_im, depth_im, rem_im, fov_up, fov_down,
                   obs_weight=1.0, use_ref=False):
  """In-place integrate of one range image into vol (dict from tsdf_new_volume).

  use_ref=False: vlo_tsdf_integrate (restatement); use_ref=True: the reference's CUDA kernel
  string compiled for the CPU (oracle/_ref/libref_tsdf.so)."""
  im_h, im_w = depth_im.shape
  trunc = voxel_size * 5  # fusion_lidar.py:31
  vd, vo, other = _tsdf_args(vol["tsdf"].shape, vol_origin, voxel_size, im_h, im_w, trunc, obs_weight,
                             fov_up, fov_down)
  color_im = np.ascontiguousarray(color_im, np.float32).reshape(-1)
  depth_im = np.ascontiguousarray(depth_im, np.float32).reshape(-1)
  rem_im = np.ascontiguousarray(rem_im, np.float32).reshape(-1)
  for k in ("tsdf", "weight", "color", "rem"):
    assert vol[k].flags["C_CONTIGUOUS"] and vol[k].dtype == np.float32
  counters = np.zeros(2, np.int64)
  if use_ref:
    lib = _lib(os.path.join(_REF_DIR, "libref_tsdf.so"))
    cam_pose = np.eye(4, dtype=np.float32).reshape(-1)
    lib.ref_tsdf_integrate(_p(vol["tsdf"], _f32p), _p(vol["weight"], _f32p), _p(vol["color"], _f32p),
                           _p(vol["rem"], _f32p), _p(vd, _f32p), _p(vo, _f32p), _p(cam_pose, _f32p),
                           _p(other, _f32p), _p(color_im, _f32p), _p(depth_im, _f32p), _p(rem_im, _f32p))
  else:
    lib = _lib(os.path.join(_HERE, "liboracle.so"))
    lib.vlo_tsdf_integrate(_p(vol["tsdf"], _f32p), _p(vol["weight"], _f32p), _p(vol["color"], _f32p),
                           _p(vol["rem"], _f32p), _p(vd, _f32p), _p(vo, _f32p), _p(other, _f32p),
                           _p(color_im, _f32p), _p(depth_im, _f32p), _p(rem_im, _f32p), _p(counters, _i64p))
  return dict(n_vis=int(counters[0]), n_written=int(counters[1]))


class RefCudaTsdf:
  """The reference's CUDA `integrate` kernel itself (oracle/_ref/libref_tsdf_cuda.so: the kernel string of
  auxiliary/fusion_lidar.py:66-229 compiled by nvcc for sm_100a with pycuda's defaults, launched with the reference's
  geometry) on torch CUDA tensors -- the checker the -m gpu tests hold the product's TSDF integration to, bit for bit.
  Mirrors TSDFVolume.__init__ / integrate (fusion_lidar.py:23-63, 252-287).  Needs a GPU; test infrastructure only."""

  def __init__(self, vol_dim, vol_origin, voxel_size, fov_up, fov_down):
    import torch
    self.dim = tuple(int(v) for v in vol_dim)
    self.origin = np.asarray(vol_origin, np.float32).copy()
    self.voxel_size, self.fov_up, self.fov_down = float(voxel_size), float(fov_up), float(fov_down)
    n = self.dim[0] * self.dim[1] * self.dim[2]
    # one element past the end belongs to the kernel's `voxel_idx > n` guard (fusion_lidar.py:92)
    self._pad = [torch.empty(n + 1, dtype=torch.float32, device="cuda") for _ in range(4)]
    self._pad[0].fill_(1.0)
    for t in self._pad[1:]:
      t.zero_()
    self.tsdf, self.weight, self.color, self.rem = (t[:n].view(self.dim) for t in self._pad)
    self.lib = _lib(os.path.join(_REF_DIR, "libref_tsdf_cuda.so"))

  def integrate(self, color_im, depth_im, rem_im, obs_weight=1.0):
    import torch
    im_h, im_w = depth_im.shape
    vd, vo, other = _tsdf_args(self.dim, self.origin, self.voxel_size, im_h, im_w, self.voxel_size * 5, obs_weight,
                               self.fov_up, self.fov_down)
    cam_pose = np.eye(4, dtype=np.float32).reshape(-1)
    color_im = np.ascontiguousarray(color_im, np.float32).reshape(-1)
    depth_im = np.ascontiguousarray(depth_im, np.float32).reshape(-1)
    rem_im = np.ascontiguousarray(rem_im, np.float32).reshape(-1)
    torch.cuda.synchronize()
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = self.lib.ref_tsdf_integrate_cuda(vp(self._pad[0]), vp(self._pad[1]), vp(self._pad[2]), vp(self._pad[3]), _p(vd, _f32p),
                                          _p(vo, _f32p), _p(cam_pose, _f32p), _p(other, _f32p), _p(color_im, _f32p),
                                          _p(depth_im, _f32p), _p(rem_im, _f32p), ctypes.c_int(im_h), ctypes.c_int(im_w))
    if rc < 0:
      raise RuntimeError("ref_tsdf_integrate_cuda failed")
    return rc


def mesh_attributes(verts_vox, color_vol, rem_vol, voxel_size, vol_origin):
  """Vertex world coords / colours / remission lookup of TSDFVolume.get_mesh,
  auxiliary/fusion_lidar.py:408-423 (numpy, verbatim order)."""
  verts_vox = np.asarray(verts_vox, np.float32)
  verts_ind = np.round(verts_vox).astype(int)
  verts = verts_vox * voxel_size + np.asarray(vol_origin, np.float32)
  rgb_vals = color_vol[verts_ind[:, 0], verts_ind[:, 1], verts_ind[:, 2]]
  rem = rem_vol[verts_ind[:, 0], verts_ind[:, 1], verts_ind[:, 2]]
  colors_b = np.floor(rgb_vals / (256 * 256))
  colors_g = np.floor((rgb_vals - colors_b * 256 * 256) / 256)
  colors_r = rgb_vals - colors_b * 256 * 256 - colors_g * 256
  colors = np.floor(np.asarray([colors_r, colors_g, colors_b])).T
  colors = colors.astype(np.int64).astype(np.uint8)
  return verts.astype(np.float32), colors, rem


def mesh_extract(tsdf, color_vol, rem_vol, voxel_size, vol_origin, level=0.0):
  """vlo_mesh_extract: iso-surface + vertex attribute lookup (restates csrc/vl_mesh.cu; the reference's
  get_mesh, auxiliary/fusion_lidar.py:403-424, delegates the surface to un-vendored scikit-image).
  Returns dict(verts f32[3T,3], faces i32[T,3], colors u8[3T,3], rem f32[3T])."""
  lib = _lib(os.path.join(_HERE, "liboracle.so"))
  lib.vlo_mesh_extract.restype = ctypes.c_longlong
  tsdf = np.ascontiguousarray(tsdf, np.float32)
  color_vol = np.ascontiguousarray(color_vol, np.float32)
  rem_vol = np.ascontiguousarray(rem_vol, np.float32)
  dx, dy, dz = tsdf.shape
  vo = np.ascontiguousarray(vol_origin, np.float32)
  args = lambda cap, v, f, c, r: lib.vlo_mesh_extract(
      _p(tsdf, _f32p), _p(color_vol, _f32p), _p(rem_vol, _f32p), ctypes.c_int(dx), ctypes.c_int(dy), ctypes.c_int(dz),
      ctypes.c_float(level), ctypes.c_float(voxel_size), _p(vo, _f32p), ctypes.c_longlong(cap), v, f, c, r)
  n = int(args(0, None, None, None, None))
  verts, faces = np.empty((3 * n, 3), np.float32), np.empty((n, 3), np.int32)
  colors, rem = np.empty((3 * n, 3), np.uint8), np.empty(3 * n, np.float32)
  if n:
    args(n, _p(verts, _f32p), _p(faces, _i32p), _p(colors, _u8p), _p(rem, _f32p))
  return dict(verts=verts, faces=faces, colors=colors, rem=rem)


def compare_numpy(source_color, target_color, source_label, target_label, source_range, target_range,
                  source_rem, target_rem, nclasses):
  """Restatement of compare(), auxiliary/laserscan.py:1181-1301, with iouEval of auxiliary/np_ioueval.py:8-70
  (pinned to the reference's own compare() by tests/golden/golden_compare_v1.npz).  Returns dict(label_diff,
  range_diff, rem_diff, conf i64[nclasses, nclasses] (rows = renumbered target label), m_iou, m_acc, mse)."""
  source_color, target_color = np.copy(source_color), np.copy(target_color)
  source_label, target_label = np.copy(source_label), np.copy(target_label)
  black = np.sum(source_color, axis=2) == 0                       # :1200
  source_label[black] = 0
  target_label[black] = 0
  target_color[black] = 0
  bg = source_label == 0                                          # :1206
  target_label[bg] = 0
  target_color[bg] = 0
  label_diff = abs(source_color - target_color)                   # :1211
  for i, value in enumerate(np.union1d(np.unique(source_label), np.unique(target_label))):   # :1214-1223
    ms, mt = source_label == value, target_label == value
    source_label[ms] = i
    target_label[mt] = i
  present = np.union1d(np.unique(source_label), np.unique(target_label))
  empty = np.isin(np.arange(nclasses), present, invert=True)      # :1225-1227
  ignore = np.arange(nclasses)[empty]
  include = np.array([n for n in range(nclasses) if n not in ignore], dtype=np.int64)
  conf = np.zeros((nclasses, nclasses), np.int64)                 # np_ioueval.py:27-45, x = target (pred), y = source
  np.add.at(conf, (target_label.reshape(-1), source_label.reshape(-1)), 1)
  m_iou, m_acc = iou_from_confusion(conf, ignore, include)
  source_range, target_range = np.copy(source_range), np.copy(target_range)
  source_range[bg] = 0                                            # :1249-1252
  target_range[bg] = 0
  range_diff = (source_range - target_range) ** 2
  mse = range_diff.sum() / range_diff.size                        # :1254
  source_rem, target_rem = np.copy(source_rem), np.copy(target_rem)
  source_rem[bg] = 0                                              # :1270-1276
  target_rem[bg] = 0
  rem_diff = (source_rem - target_rem) ** 2
  return dict(label_diff=label_diff, range_diff=range_diff, rem_diff=rem_diff, conf=conf, m_iou=m_iou, m_acc=m_acc,
              mse=mse, n_present=len(present))


def iou_from_confusion(conf, ignore, include):
  """iouEval.getStats / getIoU / getacc, auxiliary/np_ioueval.py:47-70."""
  c = conf.copy()
  c[ignore] = 0
  c[:, ignore] = 0
  tp = np.diag(c)
  fp = c.sum(axis=1) - tp
  fn = c.sum(axis=0) - tp
  union = tp + fp + fn + 1e-15
  m_iou = (tp[include] / union[include]).mean()
  m_acc = tp.sum() / (tp[include].sum() + fp[include].sum() + 1e-15)
  return m_iou, m_acc
