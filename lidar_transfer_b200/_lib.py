"""ctypes loader of libvlidar.so (the C ABI declared in include/vlidar.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing this module raises,
loudly, on first use.  Build it with `python -m lidar_transfer_b200.build`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VLIDAR_LIB", os.path.join(_HERE, "libvlidar.so"))  # VLIDAR_LIB: tuning builds only

VL_OK, VL_EINVAL, VL_ENOSPACE, VL_ECUDA, VL_EBADMESH = 0, -1, -2, -3, -4


class VlidarError(RuntimeError):
  def __init__(self, code, text):
    super().__init__("libvlidar error %d: %s" % (code, text))
    self.code = code


_c = ctypes
_vp, _i, _l, _ll, _f, _d, _sz = _c.c_void_p, _c.c_int, _c.c_long, _c.c_longlong, _c.c_float, _c.c_double, _c.c_size_t

# name -> (restype, argtypes); every symbol include/vlidar.h declares
SIGNATURES = {
    "vl_abi_version": (_i, []),
    "vl_last_error": (_c.c_char_p, []),
    "vl_device_count": (_i, []),
    "ctrace": (None, [_vp] * 6 + [_i] * 4 + [_vp] * 4),
    "vl_ctrace_status": (_i, []),
    "vl_ctrace_ids": (_i, [_vp] * 6 + [_i] * 4 + [_vp] * 5),
    "vl_bvh_blob_bytes": (_sz, [_i]),
    "vl_bvh_build": (_i, [_vp] * 4 + [_i, _i, _vp, _sz, _vp]),
    "vl_bvh_status": (_i, [_vp, _i, _vp, _vp]),
    "vl_trace": (_i, [_vp, _i, _vp, _vp, _i, _i] + [_vp] * 5 + [_i, _vp]),
    "vl_beams_bytes": (_sz, [_i, _i]),
    "vl_beams_build": (_i, [_vp, _i, _i, _vp, _sz, _i, _vp]),
    "vl_normalize_rays": (_i, [_vp, _i, _vp]),
    "vl_ctrace_normalize": (None, [_i]),
    "vl_ctrace_cache_stats": (None, [_vp, _vp]),
    "vl_ctrace_timing": (None, [_vp]),
    "vl_ctrace_wire": (None, [_i]),
    "vl_ctrace_traffic": (None, [_vp, _vp]),
    "vl_debug_pack": (_i, [_vp, _ll, _vp, _vp, _ll, _vp, _vp]),
    "vl_cast_workspace_bytes": (_sz, [_i, _i]),
    "vl_cast": (_i, [_vp] * 5 + [_i, _i, _vp, _i, _i] + [_vp] * 5 + [_i, _vp, _sz, _vp]),
    "vl_cast_status": (_i, [_vp, _vp, _vp]),
    "vl_cast_submit": (_i, [_vp] * 5 + [_i, _i, _vp, _i, _i] + [_vp] * 5 + [_i, _vp, _sz, _vp] + [_vp] * 4),
    "vl_cast_graph_create": (_i, [_vp, _vp, _i, _i] + [_vp] * 5 + [_i, _vp, _sz, _i, _vp, _vp, _vp, _vp]),
    "vl_cast_graph_launch": (_i, [_vp] * 5),
    "vl_cast_graph_destroy": (_i, [_vp]),
    "vl_ctrace_method": (None, [_i]),
    "vl_debug_mesh_scalar": (None, [_i]),
    "vl_debug_cast_cells": (None, [_i]),
    "vl_debug_cast_rearm": (None, [_i]),
    "vl_debug_cast_ctas": (None, [_i]),
    "vl_debug_cast_row_trim": (None, [_i]),
    "vl_debug_cast_split": (None, [_i]),
    "vl_debug_cast_setup_ctas": (None, [_i]),
    "vl_trace_bruteforce": (_i, [_vp] * 4 + [_i, _i, _vp, _vp, _i, _i] + [_vp] * 5 + [_i, _vp]),
    "vl_project_workspace_bytes": (_sz, [_l, _i, _i]),
    "vl_project": (_i, [_vp, _vp, _vp, _l, _d, _d, _i, _i, _i] + [_vp] * 7 + [_sz, _vp]),
    "vl_project_snap": (_i, [_vp, _vp, _vp, _l, _d, _d, _i, _i, _i, _vp, _i] + [_vp] * 7 + [_sz, _vp]),
    "vl_project_select": (_i, [_vp, _vp, _vp, _l, _d, _d, _i, _i, _i, _vp, _i, _i] + [_vp] * 7 + [_sz, _vp]),
    "vl_points_bounds": (_i, [_vp, _vp, _l, _vp, _vp]),
    "vl_reverse_project": (_i, [_vp, _vp, _vp, _i, _i, _d, _d, _vp, _vp]),
    "vl_tsdf_init": (_i, [_vp] * 4 + [_ll, _vp]),
    "vl_tsdf_integrate": (_i, [_vp] * 4 + [_i, _i, _i, _vp] + [_f] * 5 + [_vp] * 3 + [_i, _i, _vp]),
    "vl_tsdf_workspace_bytes": (_sz, [_i, _i]),
    "vl_tsdf_fresh_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "vl_debug_tsdf_shell": (None, [_i]),
    "vl_tsdf_integrate_ws": (_i, [_vp] * 4 + [_i, _i, _i, _vp] + [_f] * 5 + [_vp] * 3 + [_i, _i, _vp, _sz, _vp]),
    "vl_tsdf_init_integrate": (_i, [_vp] * 4 + [_i, _i, _i, _vp] + [_f] * 5 + [_vp] * 3 + [_i, _i, _vp, _sz, _vp]),
    "vl_tsdf_sparse_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "vl_tsdf_sparse_integrate": (_i, [_vp] * 4 + [_i, _i, _i, _vp] + [_f] * 5 + [_vp] * 3 + [_i, _i, _vp, _i, _vp, _sz, _vp]),
    "vl_tsdf_densify": (_i, [_vp] * 4 + [_i, _i, _i, _vp, _vp]),
    "vl_mesh_count_sparse": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _sz, _vp, _vp]),
    "vl_mesh_emit_sparse": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _vp, _sz, _ll, _ll, _vp] + [_vp] * 5 + [_vp]),
    "vl_mesh_workspace_bytes": (_sz, [_i, _i, _i]),
    "vl_mesh_list_bytes": (_sz, [_ll, _ll]),
    "vl_mesh_count": (_i, [_vp, _i, _i, _i, _f, _vp, _sz, _vp, _vp]),
    "vl_mesh_emit": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _sz, _ll, _ll, _vp] + [_vp] * 5 + [_vp]),
    "vl_compare_workspace_bytes": (_sz, [_i]),
    "vl_compare": (_i, [_vp] * 8 + [_i, _i] + [_vp] * 5 + [_sz, _vp]),
    "vl_compare_status": (_i, [_vp, _vp, _vp, _vp]),
    "vl_launch_count": (_ll, []),
    "vl_profile_enable": (_i, [_i]),
    "vl_profile_stage_count": (_i, []),
    "vl_profile_stage_name": (_c.c_char_p, [_i]),
    "vl_profile_collect": (_i, [_vp, _vp]),
    "vl_debug_trace_stats": (None, [_vp]),
    "vl_debug_trace_mode": (None, [_i]),
    "vl_debug_build_stop": (None, [_i]),
}

_lib = None


def lib():
  """The loaded library (ctypes releases the GIL around every call)."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise ImportError(
          "%s is missing -- the CUDA extension is required (no CPU fallback). "
          "Build it with `python -m lidar_transfer_b200.build`." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
      fn = getattr(L, name)  # AttributeError here = header/library mismatch
      fn.restype = res
      fn.argtypes = args
    if L.vl_abi_version() != 2:
      raise ImportError("libvlidar ABI version %d, expected 2" % L.vl_abi_version())
    _lib = L
  return _lib


def check(code):
  if code != VL_OK:
    raise VlidarError(code, lib().vl_last_error().decode("utf-8", "replace"))
  return code
