"""Builds lidar_transfer_b200/libvlidar.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m lidar_transfer_b200.build [--force] [--verbose]

Per-file flags: the projection and TSDF kernels state every rounding explicitly and are
compiled with -fmad=false (see the headers of vl_project.cu / vl_tsdf.cu).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvlidar.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
          "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]
SOURCES = {
    "vl_api.cu": [],
    "vl_host.cu": ["-Xcompiler", "-ffp-contract=off"],   # vl_normalize_rays: the reference's roundings, never an FMA
    "vl_bvh_build.cu": [],
    "vl_trace.cu": [],
    "vl_cast.cu": [],
    "vl_project.cu": ["-fmad=false"],
    "vl_tsdf.cu": ["-fmad=false"],
    "vl_mesh.cu": [],
    "vl_metrics.cu": [],
}
HEADERS = [os.path.join(CSRC, "vl_common.cuh"), os.path.join(HERE, "..", "include", "vlidar.h")]


def _newer(a, b):
  return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build_variant(tag, defines):
  """Tuning aid: a separate library lidar_transfer_b200/libvlidar_<tag>.so compiled with extra -D flags
  (select it at run time with VLIDAR_LIB=...)."""
  out = os.path.join(HERE, "libvlidar_%s.so" % tag)
  objdir = os.path.join(OBJ, tag)
  os.makedirs(objdir, exist_ok=True)
  objs = []
  for src, extra in SOURCES.items():
    s = os.path.join(CSRC, src)
    if not os.path.exists(s):
      continue
    o = os.path.join(objdir, src.replace(".cu", ".o"))
    objs.append(o)
    subprocess.check_call([NVCC] + COMMON + extra + ["-D" + d for d in defines] + ["-c", s, "-o", o])
  subprocess.check_call([NVCC, "-shared", "-o", out] + objs + ["-ccbin", "/usr/bin/g++"])
  return out


def build(force=False, verbose=False):
  os.makedirs(OBJ, exist_ok=True)
  objs = []
  relink = force or not os.path.exists(LIB)
  for src, extra in SOURCES.items():
    s = os.path.join(CSRC, src)
    if not os.path.exists(s):
      continue
    o = os.path.join(OBJ, src.replace(".cu", ".o"))
    objs.append(o)
    if force or _newer(s, o) or any(_newer(h, o) for h in HEADERS):
      cmd = [NVCC] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
      if verbose:
        print(" ".join(cmd))
      subprocess.check_call(cmd)
      relink = True
  if relink or any(_newer(o, LIB) for o in objs):
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++"]
    if verbose:
      print(" ".join(cmd))
    subprocess.check_call(cmd)
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
