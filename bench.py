#!/usr/bin/env python
"""bench.py -- headline benchmark of the virtual-LiDAR hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]        # our arm (N>1: launched by torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]  # the reference's CPU ray tracer

Workload ("c5-synthetic-1Mtri-64x2048", BASELINE.json configs[4] shape, the config the metric
`Mrays/s ... (64x2048 target)` and the north-star target `>= 100 Mrays/s per GPU on a 64x2048 sensor
over a ~1 M-triangle scene` are quoted on): seeded synthetic KITTI-shape scans, each with ITS OWN
~1.05 M-triangle mesh (ground grid 710 x 710 + boxes, lidar_transfer_b200/synth.py) and the HDL-64E beam
pattern 64 x 2048 = 131 072 rays.  One step = a batch of scans per GPU; per scan the closest hit of all
rays against the scan's mesh (rows (i)+(ii) of the hot path), by --method cast (default: beams indexed once
per sensor, the scan's triangles streamed through the index, vl_cast) or --method lbvh (per-scan LBVH build +
per-ray traversal); the other method is timed briefly beside it ("other_method").
Weak scaling: every rank owns its scans; there is no collective on the data path.

value         = rays cast by all ranks / device time (CUDA events), meshes and rays already resident in HBM;
                2048 scans per step and GPU, so that the timed region of the driver's 20 steps lasts > 1 s.
e2e           = the same metric through the reference-facing plugin call: auxiliary.raytracer.RayTracerCython.C_Trace
                -> extern "C" ctrace (include/vlidar.h) on PAGEABLE numpy buffers, one synchronous call per scan like
                the reference's caller (fusion_lidar.py:440-450), host<->device copies inside the timed region.
e2e_pipelined = the repo's own batch API on pinned host meshes (ScanRenderer.submit_host, 8 streams).
e2e_deform    = N=1: MultiSemLaserScan.open_multiple_scans + deform('mergemesh') on the reference's real fixture at
                config-1 size (voxel 0.05, 284 M voxels), wall time per scan.
roofline      = dominant kernel of the step (largest share of device time, CUDA events on the launching stream):
                COMPULSORY bytes (inputs once + outputs once) / mean launch duration vs MEASURED_PEAKS.json, with the
                issue-slot fraction (the binding limit of this path) beside it.
pipeline_sharded = BASELINE configs[4] as the real pipeline: a 1000-scan manifest sharded over the ranks, points in,
                results out, the mesh born on the device (scans/s at N GPUs; strong scaling of a fixed manifest).
pipeline_c1 / pipeline_c4 = the whole chain (project -> TSDF -> mesh -> cast) at BASELINE.json configs[0] / [3] size.
cpu_baseline / --impl reference = the reference's own C++ ray tracer (oracle/_ref, compiled from the
                reference sources with its shipped flags) on the same meshes, all host threads.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 64, 2048
FOV_UP, FOV_DOWN = 3.0, -25.0
N_SIDE = 710  # 2 * 709^2 = 1 005 362 ground triangles (+ boxes)
WORKLOAD = "c5-synthetic-1Mtri-64x2048"
METRIC = "Mrays/s (closest-hit ray cast of a per-scan ~1M-tri mesh, 64x2048 target)"
FIXTURE = os.path.join(ROOT, "tests", "golden", "minimal_fixture.zip")


def shared_config():
  """The keys both arms print: the workload, nothing arm-specific (the driver compares the two dicts)."""
  return {"workload": WORKLOAD, "rays_per_scan": H * W, "beams": "%dx%d HDL-64E fov +%g/%g" % (H, W, FOV_UP, FOV_DOWN),
          "tris_per_scan": "~1.05 M (ground grid %d x %d + 40 boxes, seeds 1000+k)" % (N_SIDE, N_SIDE),
          "scene_generator": "lidar_transfer_b200.synth.make_scene", "one_mesh_per_scan": True}


def _ncu(kernel):
  """Per-launch counters of `kernel` from the committed ncu --set full capture (profiles/), or {}."""
  for rnd in ("r02", "r01"):
    p = os.path.join(ROOT, "profiles", "%s_%s_ncu_full.json" % (rnd, kernel))
    try:
      return json.load(open(p))
    except (OSError, ValueError):
      continue
  return {}


def _peaks():
  """HBM roofline denominator: the driver-measured copy bandwidth in MEASURED_PEAKS.json (whatever the key is called,
  as long as it mentions hbm), else the profiling guide's fallback."""
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  try:
    d = json.load(open(p))
  except (OSError, ValueError):
    return 6650.0, "fallback (B200_PROFILING.md)"

  def walk(o, path=""):
    if isinstance(o, dict):
      for k, v in o.items():
        yield from walk(v, path + "/" + str(k))
    elif isinstance(o, (int, float)) and not isinstance(o, bool):
      yield path.lower(), float(o)
  cands = [(k, v) for k, v in walk(d) if "hbm" in k and v > 0]
  pref = [kv for kv in cands if any(t in kv[0] for t in ("gbs", "gb_s", "gbps", "gb/s", "bw", "bandwidth", "copy"))] or cands
  if not pref:
    return 6650.0, "fallback (B200_PROFILING.md; no hbm entry in MEASURED_PEAKS.json)"
  k, v = pref[0]
  if v < 100.0:   # TB/s
    v *= 1000.0
  return v, "measured (MEASURED_PEAKS.json%s)" % k


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.lines, self.proc, self.idx = [], None, gpu_index

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits", "-lms", "50"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      threading.Thread(target=self._pump, daemon=True).start()
    except OSError:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.25)
    self.proc.terminate()
    sm, smax, power, reasons = [], [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in self.lines:
      f = [x.strip() for x in ln.split(",")]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); smax.append(float(f[2]))
        power.append(float(f[3]))
      except ValueError:
        continue
      for name, val in zip(names, f[5:9]):
        if val.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
            "samples": len(sm), "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


def make_scenes(rank, n_meshes):
  from lidar_transfer_b200 import synth
  return [synth.make_scene(1000 + rank * n_meshes + k, n_side=N_SIDE) for k in range(n_meshes)]


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref), all host threads
# ------------------------------------------------------------------------------------------------
def time_reference(scenes, rays, n_calls, warmup=0):
  """Times `ctrace` of the reference C++ ray tracer (auxiliary/raytracer/RayTracer.cpp:116-124) compiled
  from the reference sources with the reference's flags (oracle/Makefile).  Returns seconds per call list."""
  from oracle import oracle as O
  kind = "reference"
  if O.have_ref("libref_raytracer.so"):
    fn = lambda sc: O.ref_ctrace(rays, np.zeros(3, np.float32), sc["verts"], sc["faces"], sc["colors"], sc["rem"], H,
                                 variant="fma")
  else:  # reference binaries not shipped: fall back to the C restatement (the oracle port)
    O.build(ref=False)
    kind = "port"
    fn = lambda sc: O.trace(rays, np.zeros(3, np.float32), sc["verts"], sc["faces"], sc["colors"], sc["rem"], H,
                            O.NORMALIZE_SSE)
  times = []
  devnull = os.open(os.devnull, os.O_WRONLY)
  saved = os.dup(1)
  os.dup2(devnull, 1)  # the reference printf()s three lines per call
  try:
    for i in range(warmup + n_calls):
      sc = scenes[i % len(scenes)]
      t0 = time.perf_counter()
      fn(sc)
      dt = time.perf_counter() - t0
      if i >= warmup:
        times.append(dt)
  finally:
    os.dup2(saved, 1)
    os.close(devnull)
  return times, kind


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  from lidar_transfer_b200.rays import create_rays
  rays = create_rays(FOV_UP, FOV_DOWN, H, W)
  scenes = make_scenes(0, 2)
  cores = os.cpu_count()
  os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core
  times, kind = time_reference(scenes, rays, n_calls=args.steps, warmup=min(args.warmup, 1))
  total = sum(times)
  value = args.steps * H * W / total / 1e6
  sample = "1 scan per step (%d tris, %d rays), %d steps; %s ctrace incl. triangle construction + BVH build" % (
      scenes[0]["faces"].shape[0], H * W, args.steps, "reference C++" if kind == "reference" else "oracle C port")
  line = {
      "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": shared_config(),
      "arm": {"scans_per_step": 1, "api": "extern \"C\" ctrace of oracle/_ref/libref_raytracer.so (reference flags), one call per scan",
              "threads": cores},
      "scans_per_s": args.steps / total,
      "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
      "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the whole chain at BASELINE.json config sizes, reported beside the headline (not part of `value`)
# ------------------------------------------------------------------------------------------------
def _fixture_scan(k):
  """Scan k of the reference's fixture (tests/golden/minimal_fixture.zip) with the `ignore` classes removed, or None."""
  import zipfile
  try:
    z = zipfile.ZipFile(FIXTURE)
    scan = np.frombuffer(z.read("minimal/sequences/00/velodyne/%06d.bin" % k), np.float32).reshape(-1, 4)
    label = np.frombuffer(z.read("minimal/sequences/00/labels/%06d.label" % k), np.uint32) & 0xFFFF
  except (OSError, KeyError):
    return None
  keep = ~np.isin(label, [0, 1])
  return scan[keep].copy(), label[keep].copy()


def _collect_stages(L):
  n_st = L.vl_profile_stage_count()
  ms_arr, cnt_arr = (ctypes.c_double * n_st)(), (ctypes.c_longlong * n_st)()
  L.vl_profile_collect(ms_arr, cnt_arr)
  return {L.vl_profile_stage_name(i).decode(): (ms_arr[i], cnt_arr[i]) for i in range(n_st) if cnt_arr[i]}


def pipeline_chain(L, dev, scans, src_fov, bnds, vox, target, reps=3):
  """points -> range image(s) -> TSDF -> iso-surface -> cast.  scans: list of (float32[N,4], uint32[N]) fused into ONE
  volume (one = mergemesh / config 1, several = the mesh adaption / config 4).  Per-stage device times from the
  library's CUDA events (second pass) and the whole chain between two events including the host's part -- allocation
  of the data-dependent outputs, the one synchronisation that reads the triangle count -- with the events off (third)."""
  import torch
  from lidar_transfer_b200 import engine
  from lidar_transfer_b200.rays import create_rays
  tH, tW, tfu, tfd = target
  dpts = [(torch.from_numpy(p[:, :3].astype(np.float64)).to(dev), torch.from_numpy(p[:, 3].copy()).to(dev),
           torch.from_numpy(l.view(np.int32).copy()).to(dev)) for p, l in scans]
  bnds = np.array(bnds, np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  beams = engine.Beams(create_rays(tfu, tfd, tH, tW), tH)
  origin = torch.zeros(3, device=dev)
  vol = engine.TsdfDevice(dim, bnds[:, 0].astype(np.float32), vox, src_fov[0], src_fov[1])
  ws, wall, stages = None, None, {}
  for rep in range(reps):   # warm-up, per-stage events on (their bookkeeping costs host time), whole-chain time with them off
    torch.cuda.synchronize()
    L.vl_profile_enable(1 if rep == 1 else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vol.reset()
    for p64, rem, lab in dpts:
      pr = engine.project(p64, rem, lab, src_fov[0], src_fov[1], H, W, workspace=ws)
      ws = pr["workspace"]
      vol.integrate(pr["proj_label"].to(torch.float32) * 65536.0, pr["range_image"], pr["proj_remissions"])
    m = vol.extract_mesh(want_norms=False, want_faces=False)   # a triangle soup: its index array is implicit
    out = engine.cast(beams, m["verts"], None, m["colors"], m["rem"], origin, zero_misses=True, check_mesh=False)
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1)
    if rep == 1:
      L.vl_profile_enable(0)
      stages = {k: round(1e3 * v[0], 1) for k, v in _collect_stages(L).items()}   # us per chain (all launches of a stage summed)
  return {"n_scans_fused": len(scans), "n_points": [int(p.shape[0]) for p, _ in scans], "voxel_size": vox,
          "volume": "%d x %d x %d" % tuple(dim), "n_voxels": int(np.prod(dim)), "n_tris": int(m["n_tris"]),
          "target": "%dx%d fov +%g/%g" % (tH, tW, tfu, tfd), "hit_fraction": float((out["range"] > 0).float().mean()),
          "stage_us": stages, "kernel_ms_per_chain": round(sum(stages.values()) / 1e3, 3), "ms_per_chain_incl_host": round(wall, 3)}


def pipelines(L, dev):
  from lidar_transfer_b200 import synth
  real = _fixture_scan(0)
  c1_scan = real if real is not None else synth.make_scan_points(1, 124668)
  c1 = pipeline_chain(L, dev, [c1_scan], (FOV_UP, FOV_DOWN), [[-50, 50], [-31, 40], [-3, 2]], 0.05, (H, W, FOV_UP, FOV_DOWN))
  c1["workload"] = "BASELINE configs[0]: %s scan 0 -> 64x2048 image -> 284 M voxels at 0.05 m (the bounds mergemesh clips to) -> mesh -> identity cast" % (
      "minimal.zip" if real is not None else "synthetic")
  # configs[3]: n_frames = 5 fused by the `mesh` adaption (one range image per scan at the source field of view, all into
  # one volume with the configuration's bounds), cast with the OS1-128 pattern.  The fixture has 3 scans, so the five are
  # synthetic KITTI-shape scans of one static world (seeds 1..5).
  c4 = pipeline_chain(L, dev, [synth.make_scan_points(s, 124668) for s in range(1, 6)], (FOV_UP, FOV_DOWN),
                      [[-50, 50], [-50, 50], [-5, 5]], 0.1, synth.SENSORS["OS1-128"])
  c4["workload"] = "BASELINE configs[3]: 5 synthetic scans fused (deform('mesh') semantics) -> 100 M voxels at 0.1 m -> mesh -> 128x2048 OS1-128 cast"
  return c1, c4


def sharded_pipeline_leg(rank, world, dev, n_manifest, barrier, reduce_max):
  """BASELINE.json configs[4] as the REAL pipeline: a manifest of n_manifest scans sharded round-robin over the ranks
  (sharding.scans_for_rank), every scan through points (pinned host, float32 as in the scan file: 1.5 MB + 1 MB up) -> range image -> 284 M-voxel TSDF
  -> mesh -> 64x2048 cast -> results (4.2 MB packed, down to pinned host memory), no collective.  The mesh is born on
  the device, so a scan moves 2.5 MB up instead of the 27 MB of the host-mesh interface.  Returns (scans, max-over-ranks ms)."""
  import torch
  from lidar_transfer_b200 import pipeline, sharding, synth
  from lidar_transfer_b200.rays import create_rays
  mine = sharding.scans_for_rank(n_manifest, rank, world)
  P = 4   # distinct point clouds per rank, cycled (scan k uses cloud k mod P)
  clouds = []
  for k in range(P):
    real = _fixture_scan(k % 3) if k < 3 else None
    pts, lab = real if real is not None else synth.make_scan_points(100 + rank * P + k, 124668)
    # the scan file's own float32 coordinates (widened to float64 on the device: exact), remissions, labels
    clouds.append((torch.from_numpy(np.ascontiguousarray(pts[:, :3], np.float32)).pin_memory(), torch.from_numpy(pts[:, 3].copy()).pin_memory(),
                   torch.from_numpy(lab.view(np.int32).copy()).pin_memory()))
  bnds = np.array([[-50, 50], [-31, 40], [-3, 2]], np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / 0.05).astype(int)
  R = H * W
  n_lanes = int(os.environ.get("VL_PIPE_LANES", "3"))   # measured: 1 / 2 / 3 / 4 / 6 scans in flight -> 1064 / 1211 / 1717 / 1676 / 447 scans per second
  pipe = pipeline.ScanPipeline(create_rays(FOV_UP, FOV_DOWN, H, W), H, FOV_UP, FOV_DOWN, bnds, 0.05, H, W, n_lanes=n_lanes, device=dev)
  for _ in pipe.run(clouds[k % P] for k in range(120)):   # warm-up: every lane has seen every cloud (grow-only buffers at their final size)
    pass
  barrier()
  t0 = time.perf_counter()
  last = None
  for _, h in pipe.run(clouds[k % P] for k in mine):
    last = h
  torch.cuda.synchronize()
  ms = 1e3 * (time.perf_counter() - t0)
  hits = float((last[24 * R:28 * R].view(torch.float32) > 0).float().mean()) if last is not None else 0.0
  barrier()
  up = sum(t.numel() * t.element_size() for t in clouds[0])
  return {"scans": n_manifest, "ms": reduce_max(ms), "h2d_bytes_per_scan": up, "d2h_bytes_per_scan": 32 * R, "hit_fraction": hits,
          "scans_this_rank": len(mine), "scans_in_flight": n_lanes}


def deform_leg(n_scans=3, reps=5):
  """The reference-shaped per-scan call the driver makes (lidar_deform.py:396-415): open_multiple_scans +
  deform('mergemesh') on the real fixture at config-1 size, files on disk -> numpy attributes.  Wall ms per scan in steady
  state: the first two passes over the three scans warm the allocator pools (the volume bounds differ per scan), the last
  three are measured."""
  import tempfile
  import zipfile
  import yaml
  if not os.path.exists(FIXTURE):
    return None
  from lidar_transfer_b200.auxiliary import laserscan as ls
  d = tempfile.mkdtemp(prefix="vl_bench_")
  zipfile.ZipFile(FIXTURE).extractall(d)
  cfg = yaml.safe_load(open(os.path.join(d, "config", "lidar_transfer.yaml")))
  src = yaml.safe_load(open(os.path.join(d, "minimal", "config.yaml")))
  seq = os.path.join(d, "minimal", "sequences", "00")
  scan_names = [os.path.join(seq, "velodyne", "%06d.bin" % k) for k in range(3)]
  label_names = [os.path.join(seq, "labels", "%06d.label" % k) for k in range(3)]
  poses = [np.eye(4) for _ in range(3)]
  t_open, t_deform = [], []
  devnull = os.open(os.devnull, os.O_WRONLY)
  saved = os.dup(1)
  os.dup2(devnull, 1)   # the classes print like the reference's
  try:
    for rep in range(reps):
      for idx in range(n_scans):
        t0 = time.perf_counter()
        scans = ls.MultiSemLaserScan(src, src, cfg["number_of_scans"], len(cfg["color_map"]), cfg["ignore"], cfg["moving"],
                                     cfg["color_map"], transformation=cfg["transformation"], preserve_float=cfg["preserve_float"],
                                     voxel_size=cfg["voxel_size"], vol_bnds=np.array(cfg["voxel_bounds"]).reshape(3, 2))
        scans.open_multiple_scans(scan_names, label_names, poses, idx)
        t1 = time.perf_counter()
        scans.deform("mergemesh", poses, idx)
        _ = scans.proj_range[0, 0] + scans.label_image[0, 0] + scans.back_points[0, 0]   # the attributes write() / compare() read
        t2 = time.perf_counter()
        if rep > 1:
          t_open.append(t1 - t0)
          t_deform.append(t2 - t1)
  finally:
    os.dup2(saved, 1)
    os.close(devnull)
  return {"api": "MultiSemLaserScan.open_multiple_scans + deform('mergemesh'), minimal.zip scans 0-2, voxel %.2f (config 1)" % cfg["voxel_size"],
          "open_ms_per_scan": round(1e3 * float(np.mean(t_open)), 2), "deform_ms_per_scan": round(1e3 * float(np.mean(t_deform)), 2),
          "rays_per_scan": H * W, "value": H * W / float(np.mean(t_open) + np.mean(t_deform)) / 1e6, "unit": "Mrays/s"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
  # stdout carries exactly ONE line (the JSON): everything libraries print there (e.g. "NCCL version ...") goes to stderr
  json_fd = os.dup(1)
  os.dup2(2, 1)
  import torch
  import torch.distributed as dist
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  numa = "off"
  if world > 1 and os.environ.get("VL_NUMA_BIND", "1") != "0":   # before any pinned allocation (first touch)
    from lidar_transfer_b200 import sharding
    numa = sharding.bind_to_gpu_numa_node(local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)

  from lidar_transfer_b200 import _lib, pipeline
  from lidar_transfer_b200.auxiliary.raytracer import RayTracerCython as rtc
  from lidar_transfer_b200.rays import create_rays
  L = _lib.lib()
  if os.environ.get("VL_CAST_REARM") is not None:   # A/B aid
    L.vl_debug_cast_rearm(int(os.environ["VL_CAST_REARM"]))
  S, K, Wm = args.scans_per_step, args.steps, args.warmup
  Se, Sp = args.e2e_scans_per_step, args.pipelined_scans_per_step
  M = min(args.distinct_meshes, S)   # distinct meshes per rank; a step cycles through them
  rays_np = create_rays(FOV_UP, FOV_DOWN, H, W)
  scenes = make_scenes(rank, M)
  n_tris = [int(sc["faces"].shape[0]) for sc in scenes]
  n_verts = [int(sc["verts"].shape[0]) for sc in scenes]
  max_f, max_v = max(n_tris), max(n_verts)
  R = H * W

  # device-resident inputs (value leg), pageable numpy inputs (e2e = plugin call), pinned host inputs (pipelined leg)
  d_scenes = [tuple(torch.from_numpy(sc[k].reshape(-1)).to(dev) for k in ("verts", "faces", "colors", "rem"))
              for sc in scenes]
  np_scenes = [tuple(np.ascontiguousarray(sc[k].reshape(-1)) for k in ("verts", "faces", "colors", "rem")) for sc in scenes]
  h_scenes = [tuple(torch.from_numpy(sc[k].reshape(-1)).pin_memory() for k in ("verts", "faces", "colors", "rem"))
              for sc in scenes]
  mesh_bytes = [sum(a.nbytes for a in ns) for ns in np_scenes]
  origin = np.zeros(3, np.float32)
  rays_flat = rays_np.reshape(-1)
  rr = pipeline.ScanRenderer(rays_np, origin, H, max_v, max_f, n_streams=args.streams, device=dev, host_io=True,
                             method=args.method)
  rr1 = None
  torch.cuda.synchronize()

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def step_device():
    for k in range(S):
      rr.submit(*d_scenes[k % M])

  def step_pipelined():
    for k in range(Sp):
      rr.submit_host(*h_scenes[k % M])

  L.vl_ctrace_method(0 if args.method == "cast" else 1)

  def step_plugin():
    """The reference caller's sequence per scan (fusion_lidar.py:440-450): zero-filled outputs, one C_Trace call."""
    for k in range(Se):
      v, f, c, r = np_scenes[k % M]
      ep, ec = np.zeros(3 * R, np.float32), np.zeros(3 * R, np.int32)
      rg, rm = np.zeros(R, np.float32), np.zeros(R, np.float32)
      rtc.C_Trace(rays_flat, origin, v, f, c, r, ep, ec, rg, rm, H, W)
    return rg

  def reduce_max(ms):
    if world > 1:
      t = torch.tensor([ms], dtype=torch.float64, device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    return ms

  def timed(step_fn, n_steps, profile, fence=None):
    """K steps bracketed by barrier + synchronize; device time from CUDA events on the current stream,
    which forks to / joins from the renderer's streams."""
    barrier()
    if profile:
      L.vl_profile_enable(1)
    launches0 = L.vl_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_steps):
      step_fn()
    (fence or (rr1 if profile else rr)).fence()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.vl_launch_count() - launches0
    stage = None
    if profile:
      L.vl_profile_enable(0)
      stage = _collect_stages(L)
    return reduce_max(ms), launches, stage

  def timed_host(step_fn, n_steps):
    """K steps of SYNCHRONOUS host calls (the plugin call returns with the results on the host): wall clock between two
    barriers; the library's kernels run on its own stream, which torch's events do not see."""
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_steps):
      step_fn()
    ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    return reduce_max(ms)

  for _ in range(Wm):  # warm-up: all legs
    step_device()
  rr.wait()
  for _ in range(min(Wm, 3)):
    step_pipelined()
  rr.wait()
  for _ in range(min(Wm, 3)):
    last_rg = step_plugin()

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  ms_dev, launches, _ = timed(step_device, K, profile=False)
  ms_e2e = timed_host(step_plugin, K)
  ms_pipe, _, _ = timed(step_pipelined, K, profile=False)
  clocks = sampler.stop() if rank == 0 else None
  phase = (ctypes.c_double * 4)()
  L.vl_ctrace_timing(phase)
  hits_c, miss_c = ctypes.c_longlong(0), ctypes.c_longlong(0)
  L.vl_ctrace_cache_stats(ctypes.byref(hits_c), ctypes.byref(miss_c))
  wire_bytes = []
  for k in range(M):   # outside the timed region: what one call per distinct mesh moves over PCIe
    v, f, c, r = np_scenes[k]
    rtc.C_Trace(rays_flat, origin, v, f, c, r, np.zeros(3 * R, np.float32), np.zeros(3 * R, np.int32), np.zeros(R, np.float32),
                np.zeros(R, np.float32), H, W)
    up, down = ctypes.c_longlong(0), ctypes.c_longlong(0)
    L.vl_ctrace_traffic(ctypes.byref(up), ctypes.byref(down))
    wire_bytes.append((up.value, down.value))

  # per-kernel durations: the distinct scans again with the library's event profiler on (events are recorded on the
  # launching stream; kept out of the headline timing because each record costs host time per launch)
  # -- on ONE stream, so that a kernel's duration is not inflated by kernels of other scans sharing the SMs
  rr1 = pipeline.ScanRenderer(rays_np, origin, H, max_v, max_f, n_streams=1, device=dev, host_io=False,
                              method=args.method, use_graph=False)

  def step_profile():
    for ds in d_scenes:
      rr1.submit(*ds)
  ms_prof, _, stage = timed(step_profile, 4, profile=True)
  n_active = n_units = 0.0
  if args.method == "cast":   # triangles that can be hit at all / work units of the last profiled scan
    info = (ctypes.c_int * 8)()
    L.vl_cast_status(ctypes.c_void_p(rr1.slots[0].blob.data_ptr()), ctypes.c_void_p(rr1.slots[0].stream.cuda_stream), info)
    n_active, n_units = float(info[1]), float(info[2])

  # the other device path on the same scans, briefly (same bits, different structure: tests/test_cast_gpu.py)
  other = "lbvh" if args.method == "cast" else "cast"
  rr2 = pipeline.ScanRenderer(rays_np, origin, H, max_v, max_f, n_streams=args.streams, device=dev, host_io=False,
                              method=other)
  So = min(S, 256)

  def step_other():
    for k in range(So):
      rr2.submit(*d_scenes[k % M])
  step_other()
  rr2.wait()
  ms_other, _, _ = timed(step_other, 3, profile=False, fence=rr2)
  other_value = So * R * world * 3 / (ms_other * 1e-3) / 1e6

  rr.wait()
  hit_frac = float((rr.slots[(S - 1) % len(rr.slots)].out["tri_id"] >= 0).float().mean().item())

  sharded = None
  if not args.no_pipeline:   # every rank takes part: the 1000-scan manifest of BASELINE.json configs[4], real pipeline
    del rr2
    rr.close()
    torch.cuda.empty_cache()
    sharded = sharded_pipeline_leg(rank, world, dev, args.manifest_scans, barrier, reduce_max)

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  value = S * R * world * K / (ms_dev * 1e-3) / 1e6
  e2e_value = Se * R * world * K / (ms_e2e * 1e-3) / 1e6
  pipe_value = Sp * R * world * K / (ms_pipe * 1e-3) / 1e6
  # bytes the plugin call moves per rank per step, as the library counted them (vl_ctrace_traffic): the staging copy packs
  # faces / colours, of the results the end points are recomputed on the host (include/vlidar.h: vl_ctrace_wire)
  h2d = sum(wire_bytes[k % M][0] for k in range(Se))
  d2h = sum(wire_bytes[k % M][1] for k in range(Se))
  peak, peak_src = _peaks()

  # roofline of the dominant kernel.  COMPULSORY bytes per launch: every input once + every output once; records and
  # work units that live between two kernels of a scan are intermediates (they stay in L2) and do not count.
  nt = float(np.mean(n_tris)); nv = float(np.mean(n_verts))
  alg_bytes = {
      "bounds": 12 * nv,
      "morton": 12 * nt + 12 * nv + 8 * nt,
      "sort_pass": 16 * nt,
      "emit_climb": 8 * nt + 12 * nt + 28 * nv + 64 * nt + 64 * 0.3 * nt,
      "top_climb": 64 * 0.004 * nt,
      "trace": 112 * nt + 12 * R + 36 * R,
      "cast_init": 8 * R,
      "cast_setup": 12 * nt + 12 * nv,                      # faces + vertices in; records / units are intermediates
      "cast_items": 16 * R + 8 * R,                         # sorted beams in, closest-hit keys out
      "cast_resolve": (8 + 4 + 16) * R + 36 * R,            # keys, slots, directions in; five outputs out
  }
  ncu_names = {"cast_setup": "cast_setup", "cast_items": "cast_items", "cast_resolve": "cast_resolve", "cast_init": "cast_init",
               "trace": "trace"}
  total_stage_ms = sum(v[0] for v in stage.values()) or 1.0
  dom = max(stage.items(), key=lambda kv: kv[1][0])[0]
  dom_ms, dom_n = stage[dom]
  achieved = alg_bytes.get(dom, 0.0) / (dom_ms / dom_n * 1e-3) / 1e9
  cap = _ncu(ncu_names.get(dom, dom))
  sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
  sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
  issue = None
  if cap.get("warp_instructions"):   # instructions issued / issue slots available during the launch (4 schedulers per SM)
    issue = float(cap["warp_instructions"]) / (dom_ms / dom_n * 1e-3 * sm_hz * sm_count * 4)
  stages_out = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share": v[0] / total_stage_ms,
                    "compulsory_GBps": alg_bytes.get(k, 0.0) / (v[0] / v[1] * 1e-3) / 1e9} for k, v in stage.items()}
  scan_bytes = 12 * nt + 28 * nv + 12 * R + 36 * R      # one scan: mesh + rays in, five outputs out
  step_gbps = scan_bytes * S * K / (ms_dev * 1e-3) / 1e9

  pipe_c1 = pipe_c4 = deform = None
  if world == 1 and not args.no_pipeline:
    torch.cuda.empty_cache()
    pipe_c1, pipe_c4 = pipelines(L, dev)
    torch.cuda.empty_cache()
    deform = deform_leg()

  # cpu baseline: the reference C++ ray tracer on a bounded sample of the same scans
  cpu = None
  if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
    n_calls = min(args.cpu_scans, M)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    times, kind = time_reference(scenes, rays_np, n_calls=n_calls, warmup=0)
    cpu_val = n_calls * R / sum(times) / 1e6
    cpu = {"value": cpu_val, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": kind,
           "sample": "%d of the step's %d distinct scans (%d tris, %d rays each), one ctrace call per scan incl. triangle "
                     "construction + BVH build; %.2f s/scan" % (n_calls, M, n_tris[0], R, sum(times) / n_calls)}

  line = {
      "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
      "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f32", "data": "synthetic",
      "config": shared_config(),
      "arm": {"method": args.method, "scans_per_step_per_gpu": S, "e2e_scans_per_step_per_gpu": Se,
              "pipelined_scans_per_step_per_gpu": Sp, "streams": args.streams, "tris_per_scan_mean": int(nt),
              "timed_region_s": {"value": ms_dev / 1e3, "e2e": ms_e2e / 1e3, "e2e_pipelined": ms_pipe / 1e3},
              "l2": "inputs larger than L2: %d distinct meshes x %.0f MB cycled per step + %.0f MB scratch per stream"
                    % (M, mesh_bytes[0] / 1e6, rr.slots[0].blob.numel() / 1e6),
              "submission": "one CUDA graph launch per scan (vl_cast_graph_launch)" if rr.use_graph else "kernel by kernel",
              "normalize": "ray directions normalised once per sensor on the host with the reference's rsqrtps + Newton step (vl_normalize_rays)",
              "beam_index": "built once per sensor outside the timed region (vl_beams_build, ~40 us)" if args.method == "cast" else None},
      "scans_per_s": S * world * K / (ms_dev * 1e-3),
      "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
              "scans_per_s": Se * world * K / (ms_e2e * 1e-3), "ms_per_step": ms_e2e / K, "ms_per_scan": ms_e2e / K / Se,
              "api": "auxiliary.raytracer.RayTracerCython.C_Trace -> extern \"C\" ctrace (include/vlidar.h) on pageable numpy "
                     "buffers, one synchronous call per scan, outputs zero-filled by the caller (fusion_lidar.py:440-450)",
              "timer": "host wall clock around synchronous calls, max over ranks",
              "caller_bytes_per_scan": {"mesh_in": mesh_bytes[0], "results_out": 32 * R},
              "wire": "staging copy packs faces (3 x 21 bit) and colours (3 x u8); end points recomputed on the host from the range",
              "last_call_phase_ms": {"beam_index_rebuild": phase[0], "stage_raycompare_h2d_issue": phase[1], "cast_and_d2h_wait": phase[2],
                                     "merge_hits": phase[3]},
              "beam_cache": {"hits": hits_c.value, "rebuilds": miss_c.value},
              "hit_fraction": float((last_rg > 0).mean())},
      "e2e_pipelined": {"value": pipe_value, "unit": "Mrays/s", "scans_per_s": Sp * world * K / (ms_pipe * 1e-3), "ms_per_step": ms_pipe / K,
                        "h2d_bytes_per_step": sum(mesh_bytes[k % M] for k in range(Sp)) * world, "d2h_bytes_per_step": Sp * R * 36 * world,
                        "api": "ScanRenderer.submit_host (pinned host mesh -> H2D -> %s -> D2H of 5 outputs), %d streams"
                               % ("vl_cast" if args.method == "cast" else "vl_bvh_build -> vl_trace", args.streams)},
      "e2e_deform": deform,
      "pipeline_sharded": None if sharded is None else {
          "workload": "BASELINE configs[4] as the real pipeline: %d scans sharded round-robin over %d GPU(s), each: points "
                      "(pinned host) -> 64x2048 image -> 284 M-voxel TSDF -> mesh -> 64x2048 cast -> results to pinned host; "
                      "no collective" % (sharded["scans"], world),
          "scans_per_s": sharded["scans"] / (sharded["ms"] * 1e-3), "value": sharded["scans"] * R / (sharded["ms"] * 1e-3) / 1e6,
          "unit": "Mrays/s", "ms_total": sharded["ms"], "ms_per_scan_per_gpu": sharded["ms"] / max(1, sharded["scans_this_rank"]),
          "h2d_bytes_per_scan": sharded["h2d_bytes_per_scan"], "d2h_bytes_per_scan": sharded["d2h_bytes_per_scan"],
          "scaling": "strong (fixed manifest)", "hit_fraction": sharded["hit_fraction"]},
      "gpu_launches": int(launches),
      "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                   "frac": achieved / peak, "traffic": cap.get("dram_bytes_per_launch"), "peak_source": peak_src,
                   "bytes": "compulsory: inputs once + outputs once (faces + vertices for cast_setup); intermediates between kernels excluded",
                   "alg_bytes_per_launch": alg_bytes.get(dom, 0.0), "ms_per_launch": dom_ms / dom_n,
                   "issue_slot_frac": issue, "warp_instructions_per_launch": cap.get("warp_instructions"),
                   "limit": "issue-bound: the reference's Moller-Trumbore arithmetic with every rounding explicit (no FMA contraction)",
                   "whole_step": {"compulsory_bytes_per_scan": scan_bytes, "GBps": step_gbps, "frac": step_gbps / peak},
                   "tris_that_can_be_hit": n_active, "work_units": n_units},
      "stages": stages_out,
      "other_method": {"method": other, "value": other_value, "unit": "Mrays/s", "ms_per_step": ms_other / 3, "scans_per_step": So},
      "pipeline_c1": pipe_c1,
      "pipeline_c4": pipe_c4,
      "cpu_baseline": cpu,
      "numa": numa,
      "clocks": clocks,
      "hit_fraction": hit_frac,
  }
  os.write(json_fd, (json.dumps(line) + "\n").encode())
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="native", choices=["native", "reference"])
  ap.add_argument("--scans-per-step", type=int, default=2048, help="device-resident scans per step and GPU (value leg)")
  ap.add_argument("--e2e-scans-per-step", type=int, default=48, help="plugin calls per step and GPU (e2e leg)")
  ap.add_argument("--pipelined-scans-per-step", type=int, default=96, help="ScanRenderer.submit_host scans per step and GPU")
  ap.add_argument("--distinct-meshes", type=int, default=8)
  ap.add_argument("--streams", type=int, default=8)
  ap.add_argument("--method", default="cast", choices=["cast", "lbvh"])
  ap.add_argument("--cpu-scans", type=int, default=8, help="scans timed for cpu_baseline (about 1.2 s each)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-pipeline", action="store_true", help="skip the config-1 / config-4 chains, the deform leg and the sharded pipeline")
  ap.add_argument("--manifest-scans", type=int, default=1000, help="scans of the sharded real-pipeline leg (all ranks together)")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
  if args.impl == "reference":
    run_reference(args)
  else:
    run_native(args)


if __name__ == "__main__":
  main()
