"""Host-side logic that needs no GPU: the C ABI library loads and exports every symbol include/vlidar.h declares,
the reference-shaped classes keep the reference's host semantics (checked against goldens produced by the
reference's own Python), scan sharding across ranks (gloo, world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")


@pytest.fixture(scope="module")
def G():
  return np.load(GOLDEN)


def _declared_symbols():
  text = open(os.path.join(ROOT, "include", "vlidar.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(ctrace|vl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(vl):
  import ctypes
  from lidar_transfer_b200 import _lib
  names = _declared_symbols()
  assert "ctrace" in names and "vl_trace" in names and "vl_bvh_build" in names and len(names) >= 25
  raw = ctypes.CDLL(_lib.LIB_PATH)
  for n in names:
    assert hasattr(raw, n), "libvlidar.so does not export %s" % n
  assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
  assert vl.vl_abi_version() == 2
  assert vl.vl_bvh_blob_bytes(0) >= 256 and vl.vl_bvh_blob_bytes(1000) > 112 * 1000
  assert vl.vl_profile_stage_count() >= 5


def test_header_is_plain_c_and_cxx(tmp_path):
  """include/vlidar.h is what a cgo / Cython / ctypes-generator binding would include: it has to compile as C99 (no C++
  types, no default arguments) and as C++."""
  import shutil, subprocess
  src = tmp_path / "hdr_check.c"
  src.write_text('#include "vlidar.h"\nint main(void) { return vl_abi_version() < 0; }\n')
  inc = os.path.join(ROOT, "include")
  if shutil.which("gcc"):
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)
  if shutil.which("g++"):
    subprocess.run(["g++", "-std=c++14", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c++", str(src)], check=True)


def test_workspace_queries_need_no_gpu(vl):
  """The size queries are pure host arithmetic (callers allocate before the first launch)."""
  n_vox = 2000 * 1420 * 100
  mesh = vl.vl_mesh_workspace_bytes(2000, 1420, 100)
  assert n_vox // 8 <= mesh < n_vox // 8 + (1 << 23)                      # one bit per voxel + unit counters
  assert vl.vl_mesh_list_bytes(4330960, 2240000) >= 8 * 2240000 + 4 * (4330960 // 256)
  col = vl.vl_tsdf_workspace_bytes(2000, 1420)
  fresh = vl.vl_tsdf_fresh_workspace_bytes(2000, 1420, 64, 2048)
  assert col >= 4 * 2000 * 1420 and fresh >= col + 8 * 64 * 2048
  assert vl.vl_project_workspace_bytes(124668, 64, 2048) >= 8 * 64 * 2048


def test_no_cuda_means_loud_failure_not_fallback(vl):
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  from lidar_transfer_b200 import engine, _lib
  with pytest.raises(RuntimeError):
    engine.require_cuda()
  with pytest.raises(RuntimeError):
    engine.Bvh(np.zeros((3, 3), np.float32), np.zeros((1, 3), np.int32), np.zeros((3, 3), np.int32), np.zeros(3, np.float32))
  # the reference-signature host entry point reports the missing device instead of computing on the CPU
  z = np.zeros(3, np.float32)
  with pytest.raises(_lib.VlidarError):
    _lib.check(vl.vl_ctrace_ids(z.ctypes.data, z.ctypes.data, z.ctypes.data, np.zeros(3, np.int32).ctypes.data,
                                np.zeros(3, np.int32).ctypes.data, z.ctypes.data, 1, 1, 1, 1, z.ctypes.data,
                                np.zeros(3, np.int32).ctypes.data, z.ctypes.data, z.ctypes.data, None))


def test_product_never_imports_the_oracle():
  for dirpath, _, files in os.walk(os.path.join(ROOT, "lidar_transfer_b200")):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        text = open(os.path.join(dirpath, f)).read()
        assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def _color_map():
  return {0: [0, 0, 0], 1: [0, 0, 255], 10: [245, 150, 100], 40: [255, 0, 255], 44: [255, 150, 255], 48: [75, 0, 75],
          50: [0, 200, 255], 51: [50, 120, 255], 70: [0, 175, 0], 71: [0, 60, 135], 72: [80, 240, 150],
          80: [150, 240, 255], 81: [0, 0, 255], 99: [255, 255, 50], 252: [245, 150, 100], 259: [255, 0, 0]}


def test_old_projection_and_pose_round_trip_match_reference(G):
  """LaserScan.do_range_projection / remove_classes / apply_pose on the host vs the reference's own Python."""
  from lidar_transfer_b200.auxiliary.laserscan import SemLaserScan
  lut = {k: v for k, v in _color_map().items()}
  for k in np.unique(G["scan_label"]):
    lut.setdefault(int(k), [1, 2, 3])
  s = SemLaserScan(64, 512, 20, color_dict=lut)
  s.points, s.remissions, s.label = G["scan_points_f32"].copy(), G["scan_rem"].copy(), G["scan_label"].copy()
  s.colorize()
  s.remove_classes([0, 1])
  s.do_range_projection(3.0, -25.0, remove=True)
  s.do_label_projection()
  assert np.array_equal(s.proj_range.view(np.int32), G["oldproj_range"].view(np.int32))
  assert np.array_equal(s.proj_label, G["oldproj_label"])
  assert np.array_equal(s.proj_remissions.view(np.int32), G["oldproj_rem"].view(np.int32))
  # pose round trip (open_multiple_scans :812-815 + deform :949) reproduces the float64 points fed to the projection
  s2 = SemLaserScan(64, 512, 20, color_dict=lut)
  s2.points, s2.remissions, s2.label = G["scan_points_f32"].copy(), G["scan_rem"].copy(), G["scan_label"].copy()
  s2.colorize()
  s2.pose = G["scan_pose"]
  s2.apply_pose()
  s2.remove_classes([0, 1])
  s2.apply_inv_pose()
  assert s2.points.dtype == np.float64 and np.array_equal(s2.points, G["proj_src_points_f64"])
  assert np.array_equal(s2.label, G["proj_src_label_in"])


def test_reverse_projection_matches_reference(G):
  from lidar_transfer_b200.auxiliary.laserscan import LaserScan
  for tag in ("src", "tgt"):
    fu, fd, H, W = G["proj_%s_args" % tag]
    H, W = int(H), int(W)
    s = LaserScan(H, W)
    s.range_image = G["proj_%s_range" % tag]
    idx = G["proj_%s_index" % tag]
    pts = G["proj_%s_kept_points" % tag]
    w = pts[idx]
    depth = np.linalg.norm(w, 2, axis=2)
    s.proj_x_float = 0.5 * (-np.arctan2(w[..., 1], w[..., 0]) / np.pi + 1.0) * W
    s.proj_y_float = (1.0 - (np.arcsin(w[..., 2] / depth) + abs(fd / 180 * np.pi)) / (abs(fd / 180 * np.pi) + abs(fu / 180 * np.pi))) * H
    s.proj_x, s.proj_y = s._clamp(s.proj_x_float, s.proj_y_float)
    for pf in (False, True):
      s.do_reverse_projection_new(fu, fd, preserve_float=pf, host=True)
      assert np.allclose(s.back_points, G["proj_%s_back_%d" % (tag, int(pf))], rtol=0, atol=1e-9), (tag, pf)


def test_write_produces_the_reference_bytes(G, tmp_path):
  from lidar_transfer_b200.auxiliary.laserscan import MultiSemLaserScan
  ms = MultiSemLaserScan.__new__(MultiSemLaserScan)
  ms.adaption = "mergemesh"
  ms.back_points = G["glue_endpoints"].copy()
  ms.label_image = G["glue_ray_colors"][:, 2].reshape(8, 32).copy()
  ms.proj_remissions = G["glue_rem_image"].copy()
  os.makedirs(tmp_path / "velodyne"); os.makedirs(tmp_path / "labels")
  ms.write(str(tmp_path), 7)
  assert np.array_equal(np.fromfile(tmp_path / "velodyne" / "000007.bin", np.uint8), G["write_bin"])
  assert np.array_equal(np.fromfile(tmp_path / "labels" / "000007.label", np.uint8), G["write_label"])
  # the same bytes (pinned to the reference's own write() by the golden) through the writer thread
  from lidar_transfer_b200.scanio import AsyncScanWriter
  with AsyncScanWriter(str(tmp_path / "async")) as w:
    ms.write(str(tmp_path / "async"), 7, writer=w)
  assert np.array_equal(np.fromfile(tmp_path / "async" / "velodyne" / "000007.bin", np.uint8), G["write_bin"])
  assert np.array_equal(np.fromfile(tmp_path / "async" / "labels" / "000007.label", np.uint8), G["write_label"])


def test_meshwrite_format(tmp_path):
  from lidar_transfer_b200.auxiliary.fusion_lidar import meshwrite
  v = np.array([[0.5, 1.25, -2.0], [1, 2, 3], [4, 5, 6.1234567]], np.float32)
  f = np.array([[0, 1, 2]], np.int32)
  c = np.array([[255, 0, 10.9], [1, 2, 3], [4, 5, 6]])
  meshwrite(str(tmp_path / "m.ply"), v, f, v, c)
  lines = open(tmp_path / "m.ply").read().splitlines()
  assert lines[0] == "ply" and lines[2] == "element vertex 3" and lines[12] == "element face 1" and lines[14] == "end_header"
  assert lines[15] == "0.500000 1.250000 -2.000000 0.500000 1.250000 -2.000000 255 0 10"
  assert lines[17].startswith("4.000000 5.000000 6.123456 ") and lines[18] == "3 0 1 2"


def test_iou_eval_known_answer():
  """auxiliary/np_ioueval.py:73-95: two overlapping squares."""
  from lidar_transfer_b200.auxiliary.np_ioueval import iouEval
  lbl = np.zeros((7, 7), dtype=np.int64); lbl[1:4, 1:4] = 1
  prd = np.zeros((7, 7), dtype=np.int64); prd[2:5, 2:5] = 1
  ev = iouEval(2, [])
  ev.addBatch(prd, lbl)
  m_iou, iou = ev.getIoU()
  assert np.isclose(iou[1], 4 / 14) and np.isclose(iou[0], 35 / 45) and np.isclose(m_iou, (4 / 14 + 35 / 45) / 2)
  assert np.isclose(ev.getacc(), 39 / 49)


def test_scan_sharding_two_ranks_gloo(tmp_path):
  """One process per GPU, scans partitioned with no collective on the data path: world_size 2 over gloo."""
  script = tmp_path / "shard.py"
  script.write_text('''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from lidar_transfer_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = sharding.scans_for_rank(23, rank, world)
t = torch.zeros(23, dtype=torch.int64)
t[mine] = rank + 1
dist.all_reduce(t)
per_rank = sharding.gather_stats({"rank": rank, "n": len(mine), "ms": 10.0 * (rank + 1)})
if rank == 0:
  print(json.dumps({"cover": t.tolist(), "stats": per_rank, "agg": sharding.aggregate(per_rank, rays_per_scan=100)}))
dist.destroy_process_group()
''' % ROOT)
  out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=300)
  assert out.returncode == 0, out.stderr[-2000:]
  import json
  line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
  d = json.loads(line)
  assert sorted(set(d["cover"])) == [1, 2] and d["cover"].count(1) == 12 and d["cover"].count(2) == 11
  assert [s["n"] for s in d["stats"]] == [12, 11]
  assert d["agg"]["scans"] == 23 and d["agg"]["ms"] == 20.0 and np.isclose(d["agg"]["mrays_per_s"], 23 * 100 / 20e-3 / 1e6)


def test_cpulist_and_measured_peak_lookup(tmp_path, monkeypatch):
  """Small host helpers: sysfs cpulist parsing for the per-GPU NUMA binding, and bench.py's tolerant lookup of the
  driver-written HBM peak (MEASURED_PEAKS.json: key names are not under this repo's control)."""
  import importlib.util
  import json
  from lidar_transfer_b200 import sharding
  assert sharding._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and sharding._cpulist("") == set()
  assert isinstance(sharding.bind_to_gpu_numa_node(0), str)   # never raises, with or without a GPU / sysfs
  spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
  bench = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(bench)
  monkeypatch.setattr(bench, "ROOT", str(tmp_path))
  assert bench._peaks()[0] == 6650.0 and "fallback" in bench._peaks()[1]
  for content, want in (({"hbm_gbs": 6552.0, "bf16_tflops": 1600.0}, 6552.0), ({"hbm": {"copy_GBps": 6400}}, 6400.0),
                        ({"hbm_tbs": 6.5}, 6500.0), ({"bf16": 1.0}, 6650.0)):
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(content))
    assert bench._peaks()[0] == want, content


def test_host_normaliser_is_the_reference_normalize_bit_for_bit(vl, oracle):
  """vl_normalize_rays (product, host) == vlo_normalize in SSE mode (oracle; pinned bit-exact to the reference's
  compiled Vector3.h:73-89 through the trace goldens) on beam grids, random directions of every magnitude, zero,
  denormal, inf and NaN components -- so the device's triangle test sees the reference's own unit vectors."""
  import ctypes
  from lidar_transfer_b200.rays import create_rays
  rng = np.random.default_rng(5)
  sets = [create_rays(3.0, -25.0, 64, 2048), create_rays(22.5, -22.5, 128, 2048), create_rays(10.67, -30.67, 32, 1024),
          (rng.normal(size=(100000, 3)) * 10.0 ** rng.uniform(-20, 18, (100000, 1))).astype(np.float32),
          np.array([[0, 0, 0], [0, 0, 1], [1e-30, 0, 0], [1e-42, 1e-43, 0], [np.inf, 1, 0], [np.nan, 0, 1], [3e38, 3e38, 3e38],
                    [-0.0, 0.0, -2.0]], np.float32)]
  for rays in sets:
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1)
    out = np.empty_like(rays)
    assert vl.vl_normalize_rays(ctypes.c_void_p(rays.ctypes.data), rays.size // 3, ctypes.c_void_p(out.ctypes.data)) == 0
    ref = oracle.normalize_rays(rays, oracle.NORMALIZE_SSE)
    assert np.array_equal(out.view(np.int32), ref.view(np.int32))
  # in place, and the IEEE mode is a different function (<= 2 ulp away, not equal everywhere)
  rays = np.ascontiguousarray(sets[0]).reshape(-1).copy()
  ref = oracle.normalize_rays(rays, oracle.NORMALIZE_SSE)
  assert vl.vl_normalize_rays(ctypes.c_void_p(rays.ctypes.data), rays.size // 3, ctypes.c_void_p(rays.ctypes.data)) == 0
  assert np.array_equal(rays.view(np.int32), ref.view(np.int32))
  ieee = oracle.normalize_rays(sets[0], 0)
  ulp = np.abs(ieee.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
  assert 0 < ulp.max() <= 4


def test_staging_pack_loops_match_numpy():
  """The packing loops of ctrace's staging copy (faces 3 x 21 bit per 64-bit word -- AVX2 eight at a time where the CPU has
  it, SSE2 / scalar otherwise --, colours one byte per component) against numpy, for every length 0 .. 40 and a long odd
  one, aligned and unaligned destinations, and the 'does not fit' detection."""
  import ctypes
  from lidar_transfer_b200 import _lib
  vl = _lib.lib()
  rng = np.random.default_rng(3)
  for n in list(range(0, 41)) + [87381, 100003]:
    faces = rng.integers(0, 1 << 21, 3 * n).astype(np.int32)
    colors = rng.integers(0, 256, 3 * n + 5).astype(np.int32)
    for shift in (0, 1):                                      # 32-byte aligned / 8-byte aligned destination
      buf = np.zeros(n + 8, np.uint64)
      base = (-buf.ctypes.data // 8) % 4
      out = buf[base + shift: base + shift + n]
      c8 = np.zeros(colors.size + 32, np.uint8)
      seen = np.zeros(2, np.uint32)
      assert vl.vl_debug_pack(faces.ctypes.data, n, out.ctypes.data, colors.ctypes.data, colors.size, c8.ctypes.data + shift, seen.ctypes.data) == 0
      f = faces.reshape(-1, 3).astype(np.uint64)
      want = f[:, 0] | (f[:, 1] << np.uint64(21)) | (f[:, 2] << np.uint64(42))
      assert np.array_equal(out, want), (n, shift)
      assert np.array_equal(c8[shift:shift + colors.size], colors.astype(np.uint8)) and not c8[shift + colors.size:].any()
      assert seen[0] >> 21 == 0 and seen[1] >> 8 == 0
  for bad in (1 << 21, -1, (1 << 31) - 1):                     # an index that no 21-bit field holds, anywhere in the array
    for pos in (0, 7, 24, 3 * 1000 - 1):
      faces = rng.integers(0, 1 << 21, 3 * 1000).astype(np.int32)
      faces[pos] = bad
      seen = np.zeros(2, np.uint32)
      out = np.zeros(1000, np.uint64)
      vl.vl_debug_pack(faces.ctypes.data, 1000, out.ctypes.data, None, 0, None, seen.ctypes.data)
      assert seen[0] >> 21 != 0, (bad, pos)
  colors = np.array([0, 255, 256, 3], np.int32)
  seen = np.zeros(2, np.uint32)
  vl.vl_debug_pack(None, 0, None, colors.ctypes.data, 4, np.zeros(4, np.uint8).ctypes.data, seen.ctypes.data)
  assert seen[1] >> 8 != 0


def test_scan_pipeline_schedule_without_a_device():
  """The software pipelining of pipeline.ScanPipeline.run, with the device work replaced by a log: results come out in
  submission order; a lane is reused only after its previous scan has been handed out; with two or more lanes the second
  half of scan k (the one that waits for the triangle count) is issued AFTER the first half of scan k + 1, so the device
  has work while the host waits; an abandoned run leaves no state behind."""
  from lidar_transfer_b200 import pipeline

  class Ev:
    def synchronize(self):
      pass

  class Lane:
    def __init__(self, i):
      self.i, self.ctx, self.tag, self.done, self.h_out = i, None, None, Ev(), "out%d" % i

  for n_lanes in (1, 2, 3, 5):
    pipe = object.__new__(pipeline.ScanPipeline)
    pipe.lanes = [Lane(i) for i in range(n_lanes)]
    log = []

    def front(lane, tag, cloud):
      assert lane.ctx is None, "lane reused before its scan was finished"
      lane.ctx, lane.tag = ("count of", tag), tag
      log.append(("front", tag, lane.i))

    def back(lane):
      assert lane.ctx is not None
      log.append(("back", lane.tag, lane.i))
      lane.ctx = None
    pipe._front, pipe._back = front, back
    clouds = [("p%d" % k, "r", "l") for k in range(11)]
    out = list(pipe.run(clouds))
    assert [t for t, _ in out] == list(range(11)) and [h for _, h in out] == ["out%d" % (k % n_lanes) for k in range(11)]
    fronts = [e[1] for e in log if e[0] == "front"]
    backs = [e[1] for e in log if e[0] == "back"]
    assert fronts == list(range(11)) and backs == list(range(11))
    pos = {(e[0], e[1]): i for i, e in enumerate(log)}
    for k in range(10):
      if n_lanes >= 2:
        assert pos[("front", k + 1)] < pos[("back", k)], (n_lanes, k)      # the next scan's kernels are queued first
      assert pos[("back", k)] < pos[("front", k + n_lanes)] if k + n_lanes < 11 else True
    # a consumer that stops half way, then a fresh run on the same object
    gen = pipe.run(clouds)
    next(gen)
    gen.close()
    log.clear()
    out = list(pipe.run([("tag%d" % k,) + c for k, c in enumerate(clouds[:4])]))
    assert [t for t, _ in out] == ["tag0", "tag1", "tag2", "tag3"]
