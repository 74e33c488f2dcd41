#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz from the REFERENCE ITSELF (run in the authoring container,
where /root/reference exists; the result is committed, the reference cannot travel to the GPU box).

    python tests/golden/make_golden.py

What is executed to produce the vectors:
  * the reference's own Python, imported unmodified from /root/reference with stub modules for the
    packages that are absent here and that these code paths never call (imageio, skimage,
    matplotlib, mpl_toolkits) and `np.float = float` (removed from numpy >= 1.24):
      - MultiSemLaserScan.create_rays                    auxiliary/laserscan.py:1092-1119
      - SemLaserScan.do_range_projection_new('depth')    auxiliary/laserscan.py:294-391
      - SemLaserScan.do_label_projection_new             auxiliary/laserscan.py:672-676
      - LaserScan.do_range_projection (vectorised)       auxiliary/laserscan.py:202-292
      - LaserScan.do_reverse_projection_new              auxiliary/laserscan.py:475-501
      - apply_pose / apply_inv_pose / remove_classes     auxiliary/laserscan.py:98-116, 652-670
      - MultiSemLaserScan.write                          auxiliary/laserscan.py:1121-1178
      - TSDFVolume.throw_rays_at_mesh -> C_Trace glue    auxiliary/fusion_lidar.py:426-455
    `auxiliary.raytracer.RayTracerCython` is replaced by a ctypes binding with C_Trace's signature
    (RayTracerCython.pyx:15-33) onto oracle/_ref/libref_raytracer_nofma.so -- the reference C++
    ray tracer compiled from its own sources (oracle/Makefile) -- because Cython's build of the
    same three .cpp files is not run here.
  * oracle/_ref/libref_raytracer{,_nofma}.so, libref_ids_nofma.so: ctrace on seeded meshes.
  * oracle/_ref/libref_tsdf.so: the reference's CUDA `integrate` kernel string
    (auxiliary/fusion_lidar.py:70-229) compiled for the CPU.

Inputs: a 1/16 decimation of scan 000000 of the reference's fixture minimal.zip (7.8 k points,
labels, pose 1 of poses.txt) and seeded synthetic meshes (lidar_transfer_b200/synth.py).
"""
import ctypes
import io
import os
import sys
import tempfile
import types
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402  (only for the ctypes plumbing onto oracle/_ref)
from lidar_transfer_b200 import synth  # noqa: E402


def import_reference():
  if not hasattr(np, "float"):
    np.float = float
  for name in ("imageio", "skimage", "skimage.measure", "matplotlib", "matplotlib.pyplot", "mpl_toolkits",
               "mpl_toolkits.mplot3d"):
    m = types.ModuleType(name)
    sys.modules[name] = m
  sys.modules["skimage"].measure = sys.modules["skimage.measure"]
  sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
  sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
  # C_Trace with the .pyx signature, bound to the reference C++ compiled from its own sources
  rtc = types.ModuleType("auxiliary.raytracer.RayTracerCython")
  lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_raytracer_nofma.so"))

  def C_Trace(rays, origin, verts, faces, colors, rem, ray_endpoints, ray_colors, range_image, rem_image, H, W):
    for a, dt in ((rays, np.float32), (origin, np.float32), (verts, np.float32), (faces, np.int32),
                  (colors, np.int32), (rem, np.float32), (ray_endpoints, np.float32), (ray_colors, np.int32),
                  (range_image, np.float32), (rem_image, np.float32)):
      if a.dtype != dt or a.ndim != 1 or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("Buffer dtype mismatch / not contiguous")
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    n_rays, n_verts, n_faces = len(rays) // 3, len(verts) // 3, len(faces) // 3
    lib.ctrace(p(rays), p(origin), p(verts), p(faces), p(colors), p(rem), ctypes.c_int(n_rays),
               ctypes.c_int(n_verts), ctypes.c_int(n_faces), ctypes.c_int(H), p(ray_endpoints), p(ray_colors),
               p(range_image), p(rem_image))
  rtc.C_Trace = C_Trace
  pkg = types.ModuleType("auxiliary.raytracer")
  pkg.__path__ = []
  pkg.RayTracerCython = rtc
  sys.modules["auxiliary.raytracer"] = pkg
  sys.modules["auxiliary.raytracer.RayTracerCython"] = rtc
  sys.path.insert(0, REF)
  import auxiliary.laserscan as LS
  import auxiliary.fusion_lidar as FL
  return LS, FL


def load_minimal():
  import yaml
  z = zipfile.ZipFile(os.path.join(REF, "minimal.zip"))
  scan = np.frombuffer(z.read("minimal/sequences/00/velodyne/000000.bin"), np.float32).reshape(-1, 4)
  label = np.frombuffer(z.read("minimal/sequences/00/labels/000000.label"), np.uint32)
  calib = {}
  for line in z.read("minimal/sequences/00/calib.txt").decode().splitlines():
    if ":" not in line:
      continue
    key, content = line.strip().split(":")
    v = [float(x) for x in content.strip().split()]
    P = np.zeros((4, 4)); P[0, :] = v[0:4]; P[1, :] = v[4:8]; P[2, :] = v[8:12]; P[3, 3] = 1.0
    calib[key] = P
  Tr, Tr_inv = calib["Tr"], np.linalg.inv(calib["Tr"])
  poses = []
  for line in z.read("minimal/sequences/00/poses.txt").decode().splitlines():
    if not line.strip():
      continue
    v = [float(x) for x in line.strip().split()]
    P = np.zeros((4, 4)); P[0, :] = v[0:4]; P[1, :] = v[4:8]; P[2, :] = v[8:12]; P[3, 3] = 1.0
    poses.append(np.matmul(Tr_inv, np.matmul(P, Tr)))  # lidar_deform.py:71
  cfg = yaml.safe_load(open(os.path.join(REF, "config", "lidar_transfer.yaml")))
  return scan, label, poses, cfg


def main():
  LS, FL = import_reference()
  scan, label, poses, cfg = load_minimal()
  color_map = cfg["color_map"]
  G = {}

  # ---- create_rays --------------------------------------------------------------------------
  for tag, (fu, fd, H, W) in dict(a=(3.0, -25.0, 8, 32), b=(10.67, -30.67, 4, 16), c=(22.5, -22.5, 3, 5)).items():
    G["rays_%s_args" % tag] = np.array([fu, fd, H, W], np.float64)
    G["rays_%s" % tag] = LS.MultiSemLaserScan.create_rays(None, fu, fd, H, W)

  # ---- projection: decimated real scan, pose round trip like deform('mergemesh') -------------
  dec = slice(0, None, 16)
  pts_in, rem_in, lab_in = scan[dec, :3].copy(), scan[dec, 3].copy(), (label[dec] & 0xFFFF).astype(np.uint32)
  G["scan_points_f32"], G["scan_rem"], G["scan_label"] = pts_in, rem_in, lab_in
  G["scan_pose"] = poses[1]
  for tag, (fu, fd, H, W) in dict(src=(3.0, -25.0, 64, 512), tgt=(10.67, -30.67, 32, 256)).items():
    s = LS.SemLaserScan(H, W, 20, color_dict=color_map)
    s.points, s.remissions, s.label = pts_in.copy(), rem_in.copy(), lab_in.copy()
    s.colorize()
    s.pose = poses[1]
    s.apply_pose()                       # open_multiple_scans, laserscan.py:812
    s.remove_classes(cfg["ignore"])      # :815
    s.apply_inv_pose()                   # deform, :949
    G["proj_%s_args" % tag] = np.array([fu, fd, H, W], np.float64)
    G["proj_%s_points_f64" % tag] = s.points.copy()
    G["proj_%s_rem_in" % tag] = s.remissions.copy()
    G["proj_%s_label_in" % tag] = s.label.astype(np.uint32).copy()
    s.do_range_projection_new(fu, fd, remove=True)
    s.do_label_projection_new()
    G["proj_%s_n_kept" % tag] = np.array([s.points.shape[0]], np.int64)
    G["proj_%s_range" % tag] = s.range_image.copy()
    G["proj_%s_index" % tag] = s.index.copy()
    G["proj_%s_label" % tag] = s.proj_label.copy()
    G["proj_%s_rem" % tag] = s.proj_remissions.copy()
    G["proj_%s_kept_points" % tag] = s.points.copy()
    # cp path: reverse projection of the same image, both float modes (laserscan.py:475-501)
    for pf in (False, True):
      s.do_reverse_projection_new(fu, fd, preserve_float=pf)
      G["proj_%s_back_%d" % (tag, int(pf))] = s.back_points.copy()
    if tag == "src":
      proj_src = s

  # ---- the older vectorised projection used for the reference scan (lidar_deform.py:408) -----
  s = LS.SemLaserScan(64, 512, 20, color_dict=color_map)
  s.points, s.remissions, s.label = pts_in.copy(), rem_in.copy(), lab_in.copy()
  s.colorize()
  s.remove_classes(cfg["ignore"])
  s.do_range_projection(3.0, -25.0, remove=True)
  s.do_label_projection()
  G["oldproj_range"], G["oldproj_idx"], G["oldproj_label"] = s.proj_range.copy(), s.proj_idx.copy(), s.proj_label.copy()
  G["oldproj_rem"] = s.proj_remissions.copy()

  # ---- TSDF: the reference CUDA kernel string on the CPU, two integrations --------------------
  vox = 0.5
  bnds = np.array([[-20, 20], [-20, 20], [-3, 2]], np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  origin = bnds[:, 0].astype(np.float32)
  vol = O.tsdf_new_volume(dim)
  proj_label3 = np.zeros(proj_src.proj_color.shape)
  proj_label3[:, :, 0] = proj_src.proj_label          # laserscan.py:970-971
  color_im = proj_label3.astype(np.float32)
  color_im = np.floor(color_im[:, :, 0] * 256 * 256 + color_im[:, :, 1] * 256 + color_im[:, :, 2])  # fusion_lidar.py:263
  G["tsdf_args"] = np.array([vox, 3.0, -25.0], np.float64)
  G["tsdf_dim"], G["tsdf_origin"], G["tsdf_color_im"] = dim, origin, color_im.astype(np.float32)
  for rep in (1, 2):
    O.tsdf_integrate(vol, origin, vox, color_im, proj_src.range_image, proj_src.proj_remissions, 3.0, -25.0,
                     use_ref=True)
    for k in ("tsdf", "weight", "color", "rem"):
      G["tsdf_%s_%d" % (k, rep)] = vol[k].copy()

  # ---- ray tracing: the reference C++ on seeded meshes + the known-answer triangle ------------
  kat_v = np.array([[-39.5, -25.5, -1.7492892], [-39.5, -25.749289, -1.5], [-39.74929, -25.5, -1.5]], np.float32)
  kat = O.ref_ctrace(np.array([[-39.5, -25.5, -1.7]], np.float32), np.zeros(3, np.float32), kat_v,
                     np.array([[0, 1, 2]], np.int32), np.array([[0, 0, 40]] * 3, np.int32),
                     np.array([.1, .2, .3], np.float32), 1, variant="nofma")  # auxiliary/raytracing.py:230-263
  G["kat_verts"], G["kat_range"], G["kat_endpoint"] = kat_v, kat["range"], kat["endpoints"]
  for tag, (seed, n_side, n_boxes, H, W, origin3) in dict(
      s=(101, 24, 4, 8, 64, (0, 0, 0)), m=(102, 90, 12, 32, 256, (0, 0, 0)), o=(103, 40, 6, 16, 64, (0.5, -0.25, 0.3))).items():
    sc = synth.make_scene(seed, n_side=n_side, n_boxes=n_boxes)
    rays = LS.MultiSemLaserScan.create_rays(None, 3.0, -25.0, H, W)
    if tag == "o":
      rays[:W] = np.array([0, 0, 1], np.float32)  # a row of misses
    org = np.array(origin3, np.float32)
    G["trace_%s_args" % tag] = np.array([seed, n_side, n_boxes, H, W], np.int64)
    G["trace_%s_origin" % tag] = org
    G["trace_%s_rays" % tag] = rays
    nofma = O.ref_ctrace(rays, org, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, variant="nofma")
    ids = O.ref_ctrace(rays, org, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, ids=True)
    fma = O.ref_ctrace(rays, org, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, variant="fma")
    for k in ("endpoints", "endcolors", "range", "endrem"):
      G["trace_%s_%s" % (tag, k)] = nofma[k]
      assert np.array_equal(nofma[k].view(np.int32), ids[k].view(np.int32)), k
    G["trace_%s_tri_id" % tag] = ids["tri_id"]
    G["trace_%s_range_fma" % tag] = fma["range"]

  # ---- throw_rays_at_mesh glue (flatten, zero-filled outputs, reshape) -----------------------
  sc = synth.make_scene(104, n_side=30, n_boxes=4)
  tv = FL.TSDFVolume.__new__(FL.TSDFVolume)
  tv.get_mesh = lambda lut: (sc["verts"], sc["faces"], None, sc["colors"].astype(np.uint8), sc["rem"])
  rays = LS.MultiSemLaserScan.create_rays(None, 3.0, -25.0, 8, 32)
  out = tv.throw_rays_at_mesh(rays, np.zeros(3, np.float32), 8, 32, None)
  G["glue_args"] = np.array([104, 30, 4, 8, 32], np.int64)
  for name, a in zip(("endpoints", "ray_colors", "verts", "colors", "faces", "range_image", "rem_image"), out):
    G["glue_" + name] = np.asarray(a)

  # ---- write(): bytes of the .bin / .label files (laserscan.py:1121-1178) ---------------------
  ms = LS.MultiSemLaserScan.__new__(LS.MultiSemLaserScan)
  ms.adaption = "mergemesh"
  ms.back_points = out[0].copy()
  ms.label_image = out[1][:, 2].reshape(8, 32).copy()
  ms.proj_remissions = out[6].copy()
  with tempfile.TemporaryDirectory() as d:
    os.makedirs(os.path.join(d, "velodyne")); os.makedirs(os.path.join(d, "labels"))
    ms.write(d, 7)
    G["write_bin"] = np.frombuffer(open(os.path.join(d, "velodyne", "000007.bin"), "rb").read(), np.uint8)
    G["write_label"] = np.frombuffer(open(os.path.join(d, "labels", "000007.label"), "rb").read(), np.uint8)

  path = os.path.join(HERE, "golden_v1.npz")
  np.savez_compressed(path, **G)
  print("wrote %s: %d arrays, %.1f kB" % (path, len(G), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
  main()
