"""Drop-in for the reference's `auxiliary.laserscan` (auxiliary/laserscan.py): LaserScan, SemLaserScan,
MultiSemLaserScan and compare() with the attribute contract lidar_deform.py:393-452 and the visualiser read
(SURVEY.md appendix D).  Scan files, poses and class filters stay on the host in numpy like the reference;
the hot path runs on the device through libvlidar:

  do_range_projection_new('depth') + do_label_projection_new   laserscan.py:294-391, 672-676 -> vl_project
  deform('mesh' | 'mergemesh')                                  laserscan.py:863-1012        -> vl_tsdf_* , vl_mesh_*,
                                                                                               vl_bvh_build, vl_trace
  create_rays                                                   laserscan.py:1092-1119       -> rays.create_rays
  write                                                         laserscan.py:1121-1178       -> same bytes, vectorised

Host-side differences that do not change results: no per-point Python loops, no test.ply side effect unless
`MultiSemLaserScan.write_ply` is set (the reference rewrites ./test.ply on every mergemesh scan, :1010).
"""
import os

import numpy as np

from .. import engine
from .. import scanio
from ..rays import create_rays as _create_rays
from . import fusion_lidar as fl
from .np_ioueval import iouEval


class _LazyAttrs(object):
  """Attributes whose value is produced on first read.  The reference computes every derived array eagerly in numpy;
  at hundreds of scans per second most of them are never looked at (the GUI's images, the per-point colours, the
  per-pixel coordinates only the `cp` adaption needs), so they are kept as thunks -- often over tensors that are still on
  the device -- and materialise, with the reference's values, dtypes and shapes, the moment somebody reads them."""

  def _lazy_set(self, name, thunk):
    d = self.__dict__
    d.pop(name, None)
    d.setdefault("_lazy", {})[name] = thunk

  def _lazy_take(self, name):
    """The current producer of `name` (a thunk returning its value), detached from the attribute."""
    d = self.__dict__
    lazy = d.get("_lazy")
    if lazy and name in lazy:
      return lazy.pop(name)
    value = d.pop(name)
    return lambda: value

  def __getattr__(self, name):   # reached only when the normal lookup fails
    lazy = self.__dict__.get("_lazy")
    if lazy and name in lazy:
      value = lazy.pop(name)()
      self.__dict__[name] = value
      return value
    raise AttributeError("%s object has no attribute %r" % (type(self).__name__, name))

  def __setattr__(self, name, value):
    lazy = self.__dict__.get("_lazy")
    if lazy:
      lazy.pop(name, None)
    object.__setattr__(self, name, value)


def _once(fn):
  """fn() evaluated at most once."""
  box = []

  def get():
    if not box:
      box.append(fn())
    return box[0]
  return get


class LaserScan(_LazyAttrs):
  """Class that contains LaserScan with x,y,z,r"""
  EXTENSIONS_SCAN = ['.bin']

  def __init__(self, H, W, transformation=None, beam_angles=None):
    self.proj_H = H
    self.proj_W = W
    if transformation is None or (not isinstance(transformation, np.ndarray) and not transformation):
      transformation = np.eye(4)
    self.transformation = np.array(transformation).reshape(4, 4)
    self.reset()
    self.beam_angles = beam_angles
    self.pose = np.eye(4)

  def reset(self):
    H, W = self.proj_H, self.proj_W
    self.points = np.zeros((0, 3), dtype=np.float32)
    self.remissions = np.zeros((0, ), dtype=np.float32)
    self.back_points = np.zeros((0, 3), dtype=np.float32)
    # the image-sized defaults (laserscan.py:66-98) are allocated when first touched
    self._lazy_set("proj_range", lambda: np.full((H, W), -1, dtype=np.float32))
    self._lazy_set("proj_xyz", lambda: np.full((H, W, 3), -1, dtype=np.float32))
    self._lazy_set("proj_remissions", lambda: np.full((H, W), -1, dtype=np.float32))
    self._lazy_set("proj_idx", lambda: np.full((H, W), -1, dtype=np.int32))
    self.proj_x = np.zeros((0, 1), dtype=np.float32)
    self.proj_y = np.zeros((0, 1), dtype=np.float32)
    self.unproj_range = np.zeros((0, 1), dtype=np.float32)
    self._lazy_set("proj_mask", lambda: np.zeros((H, W), dtype=np.int32))
    self._proj_dev = None
    self._bnds = None

  def size(self):
    return self.points.shape[0]

  def __len__(self):
    return self.size()

  # ---- IO + rigid transforms (host, float64 like the reference) -------------------------------
  @staticmethod
  def _read_bin(filename):
    if not isinstance(filename, str):
      raise TypeError("Filename should be string type, but was {type}".format(type=str(type(filename))))
    return np.fromfile(filename, dtype=np.float32).reshape((-1, 4))

  def open_scan(self, filename, fov_up, fov_down):
    self.reset()
    if not isinstance(filename, str):
      raise TypeError("Filename should be string type, but was {type}".format(type=str(type(filename))))
    if not any(filename.endswith(ext) for ext in self.EXTENSIONS_SCAN):
      raise RuntimeError("Filename extension is not valid scan file.")
    scan = self._read_bin(filename)
    self.points = scan[:, 0:3]
    self.remissions = scan[:, 3]

  def open_scan_append(self, filename, pose, fov_up, fov_down):
    scan = self._read_bin(filename)
    hom = np.ones((scan.shape[0], 4))
    hom[:, 0:3] = scan[:, 0:3]
    t_points = np.matmul(self.transformation, np.matmul(pose.reshape((-1, 4)), hom.T)).T
    if self.points.size == 0:
      self.points, self.remissions = t_points[:, 0:3], scan[:, 3]
    else:
      self.points = np.concatenate((self.points, t_points[:, 0:3]))
      self.remissions = np.concatenate((self.remissions, scan[:, 3]))

  def apply_transformation(self, transformation):
    hom = np.ones((self.points.shape[0], 4))
    hom[:, 0:3] = self.points[:, 0:3]
    self.points = np.matmul(transformation, hom.T).T[:, 0:3]

  def apply_pose(self):
    self.apply_transformation(self.pose)

  def apply_inv_pose(self):
    self.apply_transformation(np.linalg.inv(self.pose))

  def remove_points(self, keep_index):
    self.points = self.points[keep_index]
    self.remissions = self.remissions[keep_index]
    self.label = self.label[keep_index]
    if "label_color" in self.__dict__:   # still a thunk over `label` otherwise (colorize): filtering label filters it
      self.label_color = self.label_color[keep_index]

  def _remove_points_deferred(self, keep_fn):
    """remove_points with a keep mask that is still on the device: the four per-point arrays become thunks."""
    for name in ("points", "remissions", "label"):
      old = self._lazy_take(name)
      self._lazy_set(name, (lambda old: lambda: old()[keep_fn()])(old))
    if "label_color" in self.__dict__:
      old = self._lazy_take("label_color")
      self._lazy_set("label_color", (lambda old: lambda: old()[keep_fn()])(old))

  # ---- projections ---------------------------------------------------------------------------
  def _angles(self, fov_up, fov_down, remove):
    """Shared front half of both projections (laserscan.py:207-257 / 300-345): depth, image coordinates,
    optional FOV filter.  float64 numpy, used by the host-side (comparison-only) projection."""
    fov_up = fov_up / 180.0 * np.pi
    fov_down = fov_down / 180.0 * np.pi
    fov = abs(fov_down) + abs(fov_up)
    depth = np.linalg.norm(self.points, 2, axis=1)
    keep = depth != 0
    if remove:
      depth = depth[keep]
      self.remove_points(keep)
    yaw = -np.arctan2(self.points[:, 1], self.points[:, 0])
    pitch = np.arcsin(self.points[:, 2] / depth)
    if self.beam_angles:
      ba = np.asarray(self.beam_angles)
      pitch = ba[np.abs(pitch[:, None] - ba[None, :]).argmin(axis=1)]
    proj_x = 0.5 * (yaw / np.pi + 1.0)
    proj_y = 1.0 - (pitch + abs(fov_down)) / fov
    if remove:
      keep = (proj_y >= 0) & (proj_y <= 1)
      self.remove_points(keep)
      depth, proj_y, proj_x = depth[keep], proj_y[keep], proj_x[keep]
    return depth, proj_x * self.proj_W, proj_y * self.proj_H

  def _clamp(self, proj_x, proj_y):
    px = np.maximum(0, np.minimum(self.proj_W - 1, np.floor(proj_x))).astype(np.int32)
    py = np.maximum(0, np.minimum(self.proj_H - 1, np.floor(proj_y))).astype(np.int32)
    return px, py

  def do_range_projection(self, fov_up, fov_down, remove=False):
    """The older sort-and-scatter projection (laserscan.py:202-292) lidar_deform.py:408 applies to the
    single SOURCE scan that compare() measures against.  Host numpy, like the reference: it is not on the
    synthesis path.  Far-to-near scatter: the nearest point of a pixel is written last."""
    depth, proj_x, proj_y = self._angles(fov_up, fov_down, remove)
    px, py = self._clamp(proj_x, proj_y)
    self.unproj_range = np.copy(depth)
    order = np.argsort(depth)[::-1]
    self.depth = depth[order]
    px, py = px[order], py[order]
    self.proj_range[py, px] = self.depth
    self.proj_xyz[py, px] = self.points[order]
    self.proj_remissions[py, px] = self.remissions[order]
    self.proj_idx[py, px] = np.arange(depth.shape[0])[order]
    self.proj_x, self.proj_y = px, py
    self.proj_mask = (self.proj_idx > 0).astype(np.float32)

  def do_range_projection_new(self, fov_up, fov_down, remove=False, method="depth"):
    """Nearest-point-per-pixel projection (laserscan.py:294-391) as one atomicMin scatter on the device.  The images
    stay on the device (`_proj_dev`, what deform() feeds to the TSDF integration); every host attribute the reference
    sets here is a thunk with the reference's value.  The device works in float64 -- the dtype the points have after the
    pose round trip (apply_pose / apply_inv_pose), i.e. on every path deform() takes; points still in float32 (open_scan
    without a pose applied) are promoted, where the reference would compute depth / yaw / pitch in float32: bit parity is
    stated for float64 points only (the golden vectors'), a float32 cloud can differ by an ulp of range or a pixel at a border."""
    if method not in engine.PROJECT_METHODS:
      quit()   # laserscan.py:441-442
    pts = np.ascontiguousarray(self.points, np.float64)
    out = engine.project(pts, np.ascontiguousarray(self.remissions, np.float32),
                         np.ascontiguousarray(self.label).astype(np.uint32), fov_up, fov_down, self.proj_H, self.proj_W,
                         remove=remove, beam_angles=self.beam_angles if self.beam_angles else None, want_bounds=True,
                         method=method)
    meta = out["meta"].cpu().numpy()            # one 64-byte copy: bounds of the kept points + their number
    if int(meta[48:52].view(np.int32)[0]) == 0:  # the reference fails the same way at :384 (fancy index into an empty array)
      raise IndexError("do_range_projection_new: no point left after the depth / field-of-view filters")
    self._proj_dev = out
    self._fov = (fov_up, fov_down)
    H, W = self.proj_H, self.proj_W
    if remove:
      keep_fn = _once(lambda: out["keep"].cpu().numpy())
      self._bnds = meta[:48].view(np.float64).reshape(2, 3).T.copy()   # get_bnds() of the kept points, from the device
    else:  # the reference always drops depth == 0 points (:307-309), FOV filtering is optional
      keep_fn = _once(lambda: np.linalg.norm(pts, 2, axis=1) != 0)
      self._bnds = None
    self._remove_points_deferred(keep_fn)
    if method == "depthfast":
      # :418-437: the winners go into proj_range / proj_xyz / proj_remissions / proj_idx (which start at -1, :88-100 of this
      # file), index / range_image / label images keep their initial values except range_image = proj_range (:433)
      self._lazy_set("proj_idx", lambda: out["index"].cpu().numpy())
      self._lazy_set("proj_range", lambda: out["range_image"].cpu().numpy())
      self._lazy_set("proj_remissions", lambda: out["proj_remissions"].cpu().numpy())
      self._lazy_set("range_image", lambda: self.proj_range)
      self._lazy_set("index", lambda: np.full((H, W), -1, dtype=np.int32))
      self._lazy_set("label_image", lambda: np.zeros((H, W, 1)))
      self._lazy_set("label_color_image", lambda: np.zeros((H, W, 3)))

      def proj_xyz():
        im = np.full((H, W, 3), -1, dtype=np.float32)
        mask = self.proj_idx >= 0
        im[mask] = self.points[self.proj_idx[mask]]
        return im
      self._lazy_set("proj_xyz", proj_xyz)

      def sorted_coords():   # per-POINT arrays in the reference's order (:420-421, numpy's own unstable argsort)
        depth = np.linalg.norm(self.points, 2, axis=1)
        fu, fd = fov_up / 180.0 * np.pi, fov_down / 180.0 * np.pi
        xf = 0.5 * (-np.arctan2(self.points[:, 1], self.points[:, 0]) / np.pi + 1.0) * W
        yf = (1.0 - (np.arcsin(self.points[:, 2] / depth) + abs(fd)) / (abs(fd) + abs(fu))) * H
        order = np.argsort(depth)[::-1]
        px, py = self._clamp(xf, yf)
        return xf[order], yf[order], px[order], py[order]
      sorted_coords = _once(sorted_coords)
      for k, name in enumerate(("proj_x_float", "proj_y_float", "proj_x2", "proj_y2")):
        self._lazy_set(name, (lambda k: lambda: sorted_coords()[k])(k))
      return
    self._lazy_set("index", lambda: out["index"].cpu().numpy())
    self._lazy_set("range_image", lambda: out["range_image"].cpu().numpy())
    # 'pdist' never writes proj_remissions (:392-416): it keeps the -1 of :366-367
    self._lazy_set("proj_remissions", (lambda: np.full((H, W), -1, dtype=np.float32)) if method == "pdist"
                   else (lambda: out["proj_remissions"].cpu().numpy()))
    self._lazy_set("proj_range", lambda: self.range_image)

    def label_image():
      mask = self.index >= 0
      im = np.zeros((H, W, 1))
      im[mask, 0] = self.label[self.index[mask]]
      return im

    def label_color_image():
      mask = self.index >= 0
      im = np.zeros((H, W, 3))
      im[mask] = self.label_color[self.index[mask]]
      return im
    self._lazy_set("label_image", label_image)
    self._lazy_set("label_color_image", label_color_image)
    if method == "depth":
      self._lazy_set("unproj_range", lambda: np.linalg.norm(self.points, 2, axis=1))

    # per-pixel image coordinates of the winning point (float and clamped), laserscan.py:384-388
    def coords():
      w = self.points[self.index]  # index -1 wraps to the last point, exactly like the reference's fancy index
      depth = np.linalg.norm(w, 2, axis=2)
      fu, fd = fov_up / 180.0 * np.pi, fov_down / 180.0 * np.pi
      with np.errstate(invalid="ignore", divide="ignore"):
        xf = 0.5 * (-np.arctan2(w[..., 1], w[..., 0]) / np.pi + 1.0) * W
        pitch = np.arcsin(w[..., 2] / depth)
        if self.beam_angles:  # :321-327, the H*W winners only
          ba = np.asarray(self.beam_angles, np.float64)
          pitch = ba[np.abs(pitch[..., None] - ba).argmin(axis=-1)]
        yf = (1.0 - (pitch + abs(fd)) / (abs(fd) + abs(fu))) * H
      return (xf, yf) + self._clamp(xf, yf)
    coords = _once(coords)
    for k, name in enumerate(("proj_x_float", "proj_y_float", "proj_x", "proj_y")):
      self._lazy_set(name, (lambda k: lambda: coords()[k])(k))
    if method == "pdist":   # :393-401 distance of the winner's image position from its pixel centre, 1000 where no point fell
      def dist_image():
        xf, yf, px, py = coords()
        with np.errstate(invalid="ignore"):
          d = np.sqrt((yf - (py + 0.5)) ** 2 + (xf - (px + 0.5)) ** 2)
        return np.where(self.index >= 0, d, 1000).astype(np.float32)
      self._lazy_set("dist_image", dist_image)

  def do_reverse_projection_new(self, fov_up, fov_down, preserve_float=False, host=False):
    """Pixel + depth -> xyz (laserscan.py:475-501), the `cp` adaption's back projection, on the device
    (vl_reverse_project; raises without CUDA -- nothing is selected automatically).  host=True evaluates the
    reference's float64 numpy expressions instead: the comparison copy the CPU-only tests hold against the
    reference's goldens, like do_range_projection() above."""
    if preserve_float:
      px, py = self.proj_x_float, self.proj_y_float
    else:
      px, py = self.proj_x, self.proj_y
    if not host:
      self.back_points = engine.reverse_project(self.range_image, np.asarray(px, np.float64), np.asarray(py, np.float64),
                                                fov_up, fov_down).cpu().numpy()
      return
    fov_up = fov_up / 180.0 * np.pi
    fov_down = fov_down / 180.0 * np.pi
    fov = abs(fov_down) + abs(fov_up)
    depth = self.range_image
    proj_x, proj_y = px / self.proj_W, py / self.proj_H
    yaw = (proj_x * 2 - 1.0) * np.pi
    pitch = np.pi / 2 - (1.0 * fov - proj_y * fov - abs(fov_down))
    self.back_points = np.array([depth * np.sin(pitch) * np.cos(-yaw), depth * np.sin(pitch) * np.sin(-yaw),
                                 depth * np.cos(pitch)]).transpose(1, 2, 0).reshape(-1, 3)


class SemLaserScan(LaserScan):
  """Class that contains LaserScan with x,y,z,r,label,color_label"""
  EXTENSIONS_LABEL = ['.label']

  def __init__(self, H, W, nclasses, color_dict=None, transformation=None, beam_angles=None):
    super(SemLaserScan, self).__init__(H, W, transformation, beam_angles)
    self.reset()
    self.nclasses = nclasses
    self.color_dict = color_dict
    max_key = max([key + 1 for key in color_dict] + [0])
    self.color_lut = np.zeros((max_key + 100, 3), dtype=np.float32)
    for key, value in color_dict.items():
      self.color_lut[key] = np.array(value, np.float32) / 255.0

  def reset(self):
    super(SemLaserScan, self).reset()
    self.label = np.zeros((0, ), dtype=np.uint32)
    self.label_image = np.zeros((0, ), dtype=np.uint32)
    self.label_color_image = np.zeros((0, 3), dtype=np.uint32)
    self.label_color = np.zeros((0, 3), dtype=np.float32)
    H, W = self.proj_H, self.proj_W
    zeros_label = lambda: np.zeros((H, W), dtype=np.int32)
    zeros_label.is_default = True
    self._lazy_set("proj_label", zeros_label)
    self._lazy_set("proj_color", lambda: np.zeros((H, W, 3), dtype=float))

  @staticmethod
  def _read_label(filename, extensions):
    if not isinstance(filename, str):
      raise TypeError("Filename should be string type, but was {type}".format(type=str(type(filename))))
    if not any(filename.endswith(ext) for ext in extensions):
      raise RuntimeError("Filename extension is not valid label file.")
    return np.fromfile(filename, dtype=np.uint32).reshape((-1))

  def open_label(self, filename):
    label = self._read_label(filename, self.EXTENSIONS_LABEL)
    if label.shape[0] != self.points.shape[0]:
      raise ValueError("Scan and Label don't contain same number of points")
    self.label = label & 0xFFFF  # semantic label in lower half

  def open_label_append(self, filename):
    label = self._read_label(filename, self.EXTENSIONS_LABEL)
    self.label = label if self.label.size == 0 else np.concatenate((self.label, label))

  def set_label(self, label):
    if not isinstance(label, np.ndarray):
      raise TypeError("Label should be numpy array")
    if label.shape[0] != self.points.shape[0]:
      raise ValueError("Scan and Label don't contain same number of points")
    self.label = label
    self.do_label_projection()

  def colorize(self):
    # laserscan.py:640-643; a thunk over `label`: remove_points / remove_classes filter label, which filters this
    self._lazy_set("label_color", lambda: self.color_lut[self.label].reshape((-1, 3)))

  def do_label_projection(self):
    mask = self.proj_idx >= 0
    self.proj_label[mask] = self.label[self.proj_idx[mask]]
    self.proj_color[mask] = self.color_lut[self.label[self.proj_idx[mask]]]

  def do_label_projection_new(self):
    # laserscan.py:672-676: proj_label[mask] = label[index[mask]], proj_color[mask] = color_lut[...]; on top of whatever
    # the two images hold (zeros after reset()), evaluated when read -- the label image itself is already on the device
    base_label, base_color = self._lazy_take("proj_label"), self._lazy_take("proj_color")
    dev = self._proj_dev

    def proj_label():
      if getattr(base_label, "is_default", False) and dev is not None:
        return dev["proj_label"].cpu().numpy()   # zeros + labels of the winners: what the device image holds
      im, mask = base_label(), self.index >= 0
      im[mask] = self.label[self.index[mask]]
      return im

    def proj_color():
      im, mask = base_color(), self.index >= 0
      im[mask] = self.color_lut[self.label[self.index[mask]]]
      return im
    self._lazy_set("proj_label", proj_label)
    self._lazy_set("proj_color", proj_color)

  def remove_class(self, class_index):
    self.remove_classes([class_index])

  def remove_classes(self, classes):
    keep_index = ~np.isin(self.label, np.asarray(list(classes), dtype=np.int64))
    self.remove_points(keep_index)

  def get_bnds(self):
    if getattr(self, "_bnds", None) is not None and "points" not in self.__dict__:
      return self._bnds   # the kept points are still to be filtered on the host: their bounds came back from the device
    return np.concatenate((np.amin(self.points, axis=0).reshape(3, 1), np.amax(self.points, axis=0).reshape(3, 1)), axis=1)

  def _label_map(self, sequential):
    label_map = np.full(self.proj_color.shape[:2], -1, dtype=int)
    for i, (idx, rgb) in enumerate(self.color_dict.items()):
      label_map[((self.proj_color * 255).astype(np.uint8) == rgb).all(2)] = i if sequential else idx
    return label_map

  def get_label_map(self):
    return self._label_map(True)

  def convert_color_to_label(self):
    return self._label_map(False)


class MultiSemLaserScan(_LazyAttrs):
  """Class that contains multiple LaserScans with x,y,z,r,label,color_label"""
  write_ply = False  # the reference writes ./test.ply after every mergemesh deform (laserscan.py:1010)

  def __init__(self, source_config, target_config, nscans, nclasses, ignore_classes, moving_classes, color_dict=None,
               transformation=None, preserve_float=False, voxel_size=0.1, vol_bnds=None):
    self.H = source_config["beams"]
    self.W = int(source_config["fov_hor"] / source_config["angle_res_hor"])
    self.fov_up, self.fov_down = source_config["fov_up"], source_config["fov_down"]
    self.beam_angles = sorted(source_config['beam_angles']) if source_config.get('beam_angles') else None
    self.t_H = target_config["beams"]
    self.t_W = int(target_config["fov_hor"] / target_config["angle_res_hor"])
    self.t_fov_up, self.t_fov_down = target_config["fov_up"], target_config["fov_down"]
    self.t_beam_angles = self.beam_angles  # sic: the reference reads the SOURCE config here (laserscan.py:747)
    self.nscans, self.nclasses = nscans, nclasses
    self.ignore_classes, self.moving_classes = ignore_classes, moving_classes
    self.color_dict, self.transformation = color_dict, transformation
    self.preserve_float, self.voxel_size, self.vol_bnds = preserve_float, voxel_size, vol_bnds
    self.poses = np.zeros((nscans, 4, 4), dtype=np.float32)
    self.scans = [SemLaserScan(self.H, self.W, nclasses, color_dict, transformation, self.beam_angles)
                  for _ in range(self.nscans)]
    self.reset()

  def reset(self):
    for scan in self.scans:
      scan.reset()

  def get_scan(self, idx):
    return self.scans[idx]

  def open_multiple_scans(self, scan_names, label_names, poses, idx):
    self.reset()
    if self.nscans > 1:
      n_prev = self.nscans // 2
      rel = np.arange(-n_prev, self.nscans - n_prev)
      rel = np.insert(rel[rel != 0], 0, 0)  # primary scan first (it keeps its moving classes)
    else:
      rel = np.zeros(1, dtype=int)
    for i, scan in enumerate(self.scans):
      scan_idx = idx + rel[i]
      if self.nscans > 1:
        print("Open scan %d/%d %d:%d" % (i + 1, self.nscans, rel[i], scan_idx + 1))
      scan.open_scan(scan_names[scan_idx], self.fov_up, self.fov_down)
      scan.open_label(label_names[scan_idx])
      scan.colorize()
      self.poses[i] = scan.pose = poses[scan_idx]
      scan.apply_pose()
      if i != 0:
        scan.remove_classes(self.moving_classes)
      scan.remove_classes(self.ignore_classes)

  def _merge(self, H, W, beam_angles, pose):
    m = SemLaserScan(H, W, self.nclasses, self.color_dict, self.transformation, beam_angles)
    m.reset()
    m.pose = pose
    m.points = np.concatenate([m.points] + [s.points for s in self.scans])
    m.remissions = np.concatenate([m.remissions] + [s.remissions for s in self.scans])
    m.label = np.concatenate([m.label] + [s.label for s in self.scans])
    m.colorize()   # = the concatenation of the scans' label_color (laserscan.py:944): one look-up table for all, lazily
    return m

  _rays_cache = {}

  def _cast(self, tsdf_vol, lut):
    """Ray-cast the target beam pattern against the fused volume and fill the attributes write() / compare() read.
    The per-ray results stay on the device in ONE packed buffer that comes to the host, once, when the first of
    back_points / proj_range / proj_remissions / label_image / proj_color / label_color is read; the mesh deform() also
    returns stays on the device behind LazyHostArray -- the batch driver never looks at it (lidar_deform.py:415), the
    GUI and the PLY dump get a host copy the moment they do."""
    import torch
    key = (self.t_fov_up, self.t_fov_down, self.t_H, self.t_W)
    rays = MultiSemLaserScan._rays_cache.get(key)   # the target sensor's constant (create_rays, laserscan.py:981)
    if rays is None:
      rays = MultiSemLaserScan._rays_cache[key] = self.create_rays(*key)
    origin = np.array([0, 0, 0]).astype(np.float32)
    print("Get mesh by marching cubes...")
    print("Raytracing...")
    out, m = tsdf_vol.throw_rays_at_mesh_device(rays, origin, self.t_H, self.t_W)
    tH, tW, R = self.t_H, self.t_W, self.t_H * self.t_W
    lut_np = np.asarray(lut)

    def host():   # [endpoints 3R f32 | endcolors 3R i32 | range R f32 | endrem R f32], one device -> host copy
      a = out["packed"].cpu().numpy()
      return (a[:12 * R].view(np.float32).reshape(-1, 3), a[12 * R:24 * R].view(np.int32).reshape(-1, 3),
              a[24 * R:28 * R].view(np.float32).reshape(-1, tW), a[28 * R:32 * R].view(np.float32).reshape(-1, tW))
    host = _once(host)
    self._lazy_set("back_points", lambda: host()[0])
    self._lazy_set("proj_range", lambda: host()[2])
    self._lazy_set("proj_remissions", lambda: host()[3])
    self._lazy_set("label_image", lambda: np.copy(host()[1].reshape(tH, tW, 3)[:, :, 2]))   # laserscan.py:1001-1003
    self._lazy_set("label_color", lambda: lut_np[host()[1][:, 2]])
    self._lazy_set("proj_color", lambda: lut_np[self.label_image])

    def vertex_colors(t):   # lut[colors[:, 2]] (laserscan.py:1004-1005), looked up on the device when it is needed
      if isinstance(t, str):
        return (int(m["colors"].shape[0]),) + tuple(lut_np.shape[1:]) if t == "shape" else lut_np.dtype
      return torch.from_numpy(lut_np).to(t.device)[t[:, 2].long()]
    return (fl.LazyHostArray(m["verts"]), fl.LazyHostArray(m["colors"], vertex_colors), fl.LazyHostArray(m["faces"]))

  @staticmethod
  def _integrate(tsdf_vol, scan):
    """tsdf_vol.integrate(proj_label3, proj_range, proj_remissions, eye(3)) (laserscan.py:890-897, 970-975) on the
    images the projection left on the device -- unless somebody replaced the host attributes in between."""
    dev = scan._proj_dev
    if dev is not None and not any(k in scan.__dict__ for k in ("proj_label", "proj_range", "range_image", "proj_remissions")):
      tsdf_vol.integrate_device(dev["proj_label"], dev["range_image"], dev["proj_remissions"], obs_weight=1.)
      return
    proj_label3 = np.zeros(scan.proj_color.shape)
    proj_label3[:, :, 0] = scan.proj_label
    tsdf_vol.integrate(proj_label3, scan.proj_range, scan.proj_remissions, np.eye(3), obs_weight=1.)

  def deform(self, adaption, poses, idx):
    """ Deforms laserscan with specified adaption method and transformation (laserscan.py:819-1021) """
    self.adaption = adaption
    if adaption == 'cp':  # closest point: project into the TARGET image, project back
      self.merged = m = self._merge(self.t_H, self.t_W, self.t_beam_angles, poses[idx])
      m.apply_inv_pose()
      m.do_range_projection_new(self.t_fov_up, self.t_fov_down, remove=True)
      m.do_label_projection_new()
      m.do_reverse_projection_new(self.t_fov_up, self.t_fov_down, preserve_float=self.preserve_float)
      self.back_points, self.proj_range, self.proj_remissions = m.back_points, m.proj_range, m.proj_remissions
      self.proj_color, self.label_image, self.index = m.proj_color, m.label_image, m.index
      self.label_color = m.label_color_image.reshape(-1, 3)
      return [], [], []

    elif adaption == 'mesh':  # one range image per scan, all fused into one volume
      vol_bnds = self.vol_bnds
      inv = np.linalg.inv(poses[idx])
      for i, scan in enumerate(self.scans):
        print("Create range image %d/%d" % (i + 1, self.nscans))
        scan.apply_transformation(inv)
        scan.do_range_projection_new(self.fov_up, self.fov_down, remove=True)
        scan.do_label_projection_new()
      print("Initializing voxel volume...")
      tsdf_vol = fl.TSDFVolume(vol_bnds, voxel_size=self.voxel_size, fov_up=self.fov_up, fov_down=self.fov_down)
      for i, scan in enumerate(self.scans):
        print("Fusing scan %d/%d" % (i + 1, self.nscans))
        self._integrate(tsdf_vol, scan)
      return self._cast(tsdf_vol, self.scans[0].color_lut)

    elif adaption == 'mergemesh':  # all scans merged into ONE range image (source size, TARGET fov), then fused
      self.merged = m = self._merge(self.H, self.W, self.beam_angles, poses[idx])
      m.apply_inv_pose()
      m.do_range_projection_new(self.t_fov_up, self.t_fov_down, remove=True)
      m.do_label_projection_new()
      merged_bnds = np.rint(m.get_bnds()).astype(int)
      vol_bnds = self.vol_bnds  # clipped IN PLACE like the reference (the caller's array is shared across scans)
      vol_bnds[:, 0] = np.maximum(vol_bnds[:, 0], merged_bnds[:, 0])
      vol_bnds[:, 1] = np.minimum(vol_bnds[:, 1], merged_bnds[:, 1])
      print("Initializing voxel volume...")
      tsdf_vol = fl.TSDFVolume(vol_bnds, voxel_size=self.voxel_size, fov_up=self.t_fov_up, fov_down=self.t_fov_down)
      self._integrate(tsdf_vol, m)
      print("target dim:", self.t_H, self.t_W)
      verts, colors, faces = self._cast(tsdf_vol, m.color_lut)
      if self.write_ply:
        fl.meshwrite("test.ply", verts, faces, verts, colors[..., ::-1] * 255)
      return verts, colors, faces

    elif adaption == 'catmesh':
      raise NotImplementedError("'catmesh' is a stub in the reference too (laserscan.py:1014-1016)")
    else:
      raise ValueError("Adaption method not recognized or not defined: %r" % (adaption,))

  def get_label_map(self):
    proj_color = self.label_color.reshape(self.H, self.W, 3)
    label_map = np.full(proj_color.shape[:2], -1, dtype=int)
    for i, (idx, rgb) in enumerate(self.color_dict.items()):
      label_map[(proj_color.astype(np.uint8) == rgb).all(2)] = i
    return label_map

  def create_rays(self, fov_up, fov_down, H, W):
    return _create_rays(fov_up, fov_down, H, W)

  def write(self, out_dir, idx, write_png=False, writer=None):
    """KITTI .bin (float32 x,y,z,remission) + .label (uint32) of the re-rendered scan; the reference's filter
    rules (laserscan.py:1133-1158) with array writes instead of per-point struct.pack.  writer: a
    scanio.AsyncScanWriter -- filtering and the two file writes then happen on its thread (same bytes)."""
    if write_png:
      raise NotImplementedError("write_png references an undefined name in the reference (laserscan.py:1124-1126)")
    if self.adaption == 'cp':
      back_points = self.merged.back_points.reshape(-1, 3)
      label_image = self.merged.label_image.reshape(-1)
      remissions = self.merged.proj_remissions.reshape(-1)
      valid = self.merged.index.reshape(-1,) > 0
      back_points, remissions = back_points[valid], remissions[valid]
      label_image = label_image[valid].astype(np.int32)
    else:
      back_points = self.back_points.reshape(-1, 3)
      label_image = self.label_image.reshape(-1)
      remissions = self.proj_remissions.reshape(-1)
    if writer is not None:
      writer.submit(idx, np.asarray(back_points), np.asarray(remissions), np.asarray(label_image))
      return
    rec, labels = scanio.filter_scan_for_write(back_points, remissions, label_image)
    scanio.write_scan(out_dir, idx, rec, labels)


def compare_device(scan_source, scan_target):
  """compare() on the device (engine.compare -> vl_compare): same masks, renumbering and confusion matrix; the three
  difference images come back as float32 numpy arrays, m_iou / m_acc are identical to compare()'s, MSE is accumulated
  in double (compare() sums float32 pairwise: equal to ~1e-7 relative)."""
  from .. import engine
  if scan_target.adaption == 'cp':
    target_label, target_color = scan_target.merged.proj_label, scan_target.merged.proj_color
  else:
    target_color, target_label = scan_target.proj_color, scan_target.label_image
  r = engine.compare(np.asarray(scan_source.proj_color), np.asarray(target_color), np.asarray(scan_source.proj_label),
                     np.asarray(target_label), np.asarray(scan_source.proj_range), np.asarray(scan_target.proj_range),
                     np.asarray(scan_source.proj_remissions), np.asarray(scan_target.proj_remissions), scan_source.nclasses)
  print("IoU: ", r["m_iou"])
  print("Acc: ", r["m_acc"])
  print("MSE: ", r["mse"])
  return (r["label_diff"].cpu().numpy(), r["range_diff"].cpu().numpy(), r["rem_diff"].cpu().numpy(),
          r["m_iou"], r["m_acc"], r["mse"])


def compare(scan_source, scan_target):
  """ Compare two scans by examine labels, range and remissions (laserscan.py:1181-1301): the identity
  re-render self-check.  Returns (label_diff, range_diff, remissions_diff, m_iou, m_acc, MSE).
  (Host numpy, operation for operation like the reference; compare_device() is the same on the GPU.) """
  source_color = np.copy(scan_source.proj_color)
  source_label = np.copy(scan_source.proj_label)
  if scan_target.adaption == 'cp':
    target_label = np.copy(scan_target.merged.proj_label)
    target_color = np.copy(scan_target.merged.proj_color)
  else:
    target_color = np.copy(scan_target.proj_color)
    target_label = np.copy(scan_target.label_image)
  assert source_color.size == target_color.size
  assert source_label.size == target_label.size

  # no data (black) in the source scan masks both; source background masks the target
  black = np.sum(source_color, axis=2) == 0
  source_label[black] = 0
  target_label[black] = 0
  target_color[black] = 0
  bg_label = source_label == 0
  target_label[bg_label] = 0
  target_color[bg_label] = 0
  label_diff = abs(source_color - target_color)

  # compress the labels that occur to 0..k-1 (sequentially, in place, like the reference :1217-1223)
  for i, value in enumerate(np.union1d(np.unique(source_label), np.unique(target_label))):
    mask_source, mask_target = source_label == value, target_label == value
    source_label[mask_source] = i
    target_label[mask_target] = i
  present = np.union1d(np.unique(source_label), np.unique(target_label))
  empty = np.isin(np.arange(scan_source.nclasses), present, invert=True)
  ev = iouEval(scan_source.nclasses, np.arange(scan_source.nclasses)[empty])
  ev.addBatch(target_label, source_label)
  m_iou, iou = ev.getIoU()
  print("IoU class: ", (iou * 100).astype(int))
  m_acc = ev.getacc()
  print("IoU: ", m_iou)
  print("Acc: ", m_acc)

  source_range, target_range = np.copy(scan_source.proj_range), np.copy(scan_target.proj_range)
  source_range[bg_label] = 0
  target_range[bg_label] = 0
  range_diff = (source_range - target_range) ** 2
  MSE = range_diff.sum() / range_diff.size
  print("MSE: ", MSE)

  source_rem, target_rem = np.copy(scan_source.proj_remissions), np.copy(scan_target.proj_remissions)
  source_rem[bg_label] = 0
  target_rem[bg_label] = 0
  remissions_diff = (source_rem - target_rem) ** 2
  return label_diff, range_diff, remissions_diff, m_iou, m_acc, MSE
