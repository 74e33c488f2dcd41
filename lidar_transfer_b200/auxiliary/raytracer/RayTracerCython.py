"""Drop-in for the reference's Cython module `auxiliary.raytracer.RayTracerCython`.

`C_Trace` keeps the signature and buffer contract of auxiliary/raytracer/RayTracerCython.pyx:15-33:
ten 1-D C-contiguous typed buffers (float32 / int32, anything exposing the buffer protocol) plus H, W;
wrong dtype, dimensionality or stride raises ValueError exactly where Cython's typed memoryviews
(`float[::1]`, `int[::1]`) would; `W` is accepted and ignored (the reference derives
width = n_rays / H in C, RayTracer.cpp:56); returns None; outputs are written in place for HITS ONLY.

It binds `extern "C" ctrace` of libvlidar.so (include/vlidar.h) -- the same symbol name and argument
list the .pyx declares at :5-7 -- so the call is: host buffers -> H2D -> LBVH build -> traversal -> D2H.
"""
import ctypes

import numpy as np

from ... import _lib


def _view(buf, dtype, name):
  a = np.asarray(buf) if not isinstance(buf, np.ndarray) else buf
  if a.dtype != dtype:
    raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s' (%s)" % (
        "float" if dtype == np.float32 else "int", a.dtype.name, name))
  if a.ndim != 1:
    raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d) (%s)" % (a.ndim, name))
  if not a.flags["C_CONTIGUOUS"]:
    raise ValueError("ndarray is not C-contiguous (%s)" % name)
  return a


def C_Trace(rays, origin, verts, faces, colors, rem, ray_endpoints, ray_colors, range_image, rem_image, H, W):
  f32, i32 = np.float32, np.int32
  rays, origin, verts = _view(rays, f32, "rays"), _view(origin, f32, "origin"), _view(verts, f32, "verts")
  faces, colors, rem = _view(faces, i32, "faces"), _view(colors, i32, "colors"), _view(rem, f32, "rem")
  ray_endpoints, ray_colors = _view(ray_endpoints, f32, "ray_endpoints"), _view(ray_colors, i32, "ray_colors")
  range_image, rem_image = _view(range_image, f32, "range_image"), _view(rem_image, f32, "rem_image")
  for a, name in ((ray_endpoints, "ray_endpoints"), (ray_colors, "ray_colors"), (range_image, "range_image"),
                  (rem_image, "rem_image")):
    if not a.flags["WRITEABLE"]:
      raise ValueError("buffer source array is read-only (%s)" % name)
  n_rays = len(rays) // 3
  n_verts = len(verts) // 3  # = n_colors
  n_faces = len(faces) // 3
  # the reference indexes &buf[0] of every view: an empty buffer is an IndexError there
  for a, name in ((rays, "rays"), (origin, "origin"), (verts, "verts"), (faces, "faces"), (colors, "colors"),
                  (rem, "rem"), (ray_endpoints, "ray_endpoints"), (ray_colors, "ray_colors"),
                  (range_image, "range_image"), (rem_image, "rem_image")):
    if a.shape[0] == 0:
      raise IndexError("Out of bounds on buffer access (axis 0) (%s)" % name)
  # sizes the C side trusts blindly in the reference (out-of-bounds writes there); checked here
  if len(origin) < 3 or len(colors) < 3 * n_verts or len(rem) < n_verts or len(ray_endpoints) < 3 * n_rays or \
      len(ray_colors) < 3 * n_rays or len(range_image) < n_rays or len(rem_image) < n_rays:
    raise ValueError("C_Trace: a buffer is shorter than n_rays / n_verts requires")
  L = _lib.lib()
  p = lambda a: ctypes.c_void_p(a.ctypes.data)
  L.ctrace(p(rays), p(origin), p(verts), p(faces), p(colors), p(rem), n_rays, n_verts, n_faces, int(H),
           p(ray_endpoints), p(ray_colors), p(range_image), p(rem_image))
  _lib.check(L.vl_ctrace_status())
