#!/usr/bin/env python
"""bench.py -- headline benchmark of the virtual-LiDAR hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]        # our arm (N>1: launched by torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]  # the reference's CPU ray tracer

Workload ("c5-synthetic-1Mtri-64x2048", BASELINE.json configs[4] shape, the config the metric
`Mrays/s ... (64x2048 target)` and the north-star target `>= 100 Mrays/s per GPU on a 64x2048 sensor
over a ~1 M-triangle scene` are quoted on): S seeded synthetic KITTI-shape scans per GPU per step,
each with ITS OWN ~1.0 M-triangle mesh (ground grid 710 x 710 + boxes, lidar_transfer_b200/synth.py)
and the HDL-64E beam pattern 64 x 2048 = 131 072 rays.  One step = for every scan of the batch: the closest hit
of all rays against the scan's mesh (rows (i)+(ii) of the hot path), by --method cast (default: beams indexed once
per sensor, the scan's triangles streamed through the index, vl_cast) or --method lbvh (per-scan LBVH build +
per-ray traversal, vl_bvh_build + vl_trace); the other method is timed briefly beside it ("other_method").
Weak scaling: every rank owns S scans; there is no collective on the data path.

value   = rays traced by all ranks / device time, meshes and rays already resident in HBM.
e2e     = same metric through the host-buffer path (pinned host meshes -> H2D -> cast -> D2H of the five
          per-ray outputs), copies inside the timed region.
roofline= dominant kernel of the step (largest share of device time, measured live with CUDA events
          on the launching streams): algorithmic bytes / mean launch duration vs MEASURED_PEAKS.json.
cpu_baseline / --impl reference = the reference's own C++ ray tracer (oracle/_ref, compiled from the
          reference sources with its shipped flags) on the same meshes, all host threads.
"""
import argparse
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


def _ncu_traffic(kernel):
  """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), or None."""
  p = os.path.join(ROOT, "profiles", "r01_%s_ncu_full.json" % kernel)
  try:
    return float(json.load(open(p))["dram_bytes_per_launch"])
  except (OSError, KeyError, ValueError):
    return None


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
    sm, smax, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in self.lines:
      f = [x.strip() for x in ln.split(",")]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); smax.append(float(f[2]))
      except ValueError:
        continue
      for name, val in zip(names, f[5:9]):
        if val.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
            "samples": len(sm), "reasons": sorted(reasons)}


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
      "config": {"workload": WORKLOAD, "scans_per_step": 1, "rays_per_scan": H * W,
                 "tris_per_scan": int(scenes[0]["faces"].shape[0])},
      "scans_per_s": args.steps / total,
      "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
      "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the whole chain at BASELINE.json config-1 size, reported beside the headline (not part of `value`)
# ------------------------------------------------------------------------------------------------
def pipeline_c1(L, dev):
  """points -> range image -> 284 M-voxel TSDF (voxel 0.05 m) -> iso-surface -> cast, one synthetic 124 668-point
  scan: per-stage device times from the library's CUDA events (second pass) and the whole chain between two events
  including the host's part -- allocation of the data-dependent outputs, the one synchronisation that reads the
  triangle count -- with the per-stage events off (third pass)."""
  import ctypes
  import torch
  from lidar_transfer_b200 import engine, synth
  from lidar_transfer_b200.rays import create_rays
  pts, labels = synth.make_scan_points(1, 124668)
  p64 = torch.from_numpy(pts[:, :3].astype(np.float64)).to(dev)
  rem = torch.from_numpy(pts[:, 3].copy()).to(dev)
  lab = torch.from_numpy(labels.view(np.int32)).to(dev)
  vox = 0.05
  bnds = np.array([[-50, 50], [-35.5, 35.5], [-3, 2]], np.float64)
  dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / vox).astype(int)
  rays = torch.from_numpy(create_rays(FOV_UP, FOV_DOWN, H, W)).to(dev)
  origin = torch.zeros(3, device=dev)
  beams = engine.Beams(rays, H)
  vol = engine.TsdfDevice(dim, bnds[:, 0].astype(np.float32), vox, FOV_UP, FOV_DOWN)
  ws = None
  wall = None
  for rep in range(3):   # warm-up, per-stage events on (their bookkeeping costs host time), whole-chain time with them off
    torch.cuda.synchronize()
    L.vl_profile_enable(1 if rep == 1 else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pr = engine.project(p64, rem, lab, FOV_UP, FOV_DOWN, H, W, workspace=ws)
    ws = pr["workspace"]
    vol.reset()
    vol.integrate(pr["proj_label"].to(torch.float32) * 65536.0, pr["range_image"], pr["proj_remissions"])
    m = vol.extract_mesh(want_norms=False)
    out = engine.cast(beams, m["verts"], m["faces"], m["colors"].to(torch.int32), m["rem"], origin, zero_misses=True,
                      check_mesh=False)
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1)
  L.vl_profile_enable(0)
  n_st = L.vl_profile_stage_count()
  ms_arr, cnt_arr = (ctypes.c_double * n_st)(), (ctypes.c_longlong * n_st)()
  L.vl_profile_collect(ms_arr, cnt_arr)
  stages = {L.vl_profile_stage_name(i).decode(): round(1e3 * ms_arr[i] / cnt_arr[i], 1) for i in range(n_st) if cnt_arr[i]}
  return {"workload": "c1-shape: 124668 points -> 64x2048 image -> %d x %d x %d voxels -> mesh -> 64x2048 cast" % tuple(dim),
          "n_voxels": int(np.prod(dim)), "n_tris": int(m["faces"].shape[0]),
          "hit_fraction": float((out["range"] > 0).float().mean()),
          "stage_us": stages, "kernel_ms_per_scan": round(sum(stages.values()) / 1e3, 3),
          "ms_per_scan_incl_host": round(wall, 3)}


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

  from lidar_transfer_b200 import _lib, engine, pipeline
  from lidar_transfer_b200.rays import create_rays
  L = _lib.lib()
  S, K, Wm = args.scans_per_step, args.steps, args.warmup
  M = min(args.distinct_meshes, S)   # distinct meshes per rank; a step cycles through them S / M times
  rays_np = create_rays(FOV_UP, FOV_DOWN, H, W)
  scenes = make_scenes(rank, M)
  n_tris = [int(sc["faces"].shape[0]) for sc in scenes]
  n_verts = [int(sc["verts"].shape[0]) for sc in scenes]
  max_f, max_v = max(n_tris), max(n_verts)

  # device-resident inputs (value leg) and pinned host inputs (e2e leg)
  d_scenes = [tuple(torch.from_numpy(sc[k].reshape(-1)).to(dev) for k in ("verts", "faces", "colors", "rem"))
              for sc in scenes]
  h_scenes = [tuple(torch.from_numpy(sc[k].reshape(-1)).pin_memory() for k in ("verts", "faces", "colors", "rem"))
              for sc in scenes]
  mesh_bytes = [sum(t.numel() * t.element_size() for t in hs) for hs in h_scenes]
  origin = np.zeros(3, np.float32)
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

  def step_host():
    for k in range(S):
      rr.submit_host(*h_scenes[k % M])

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
      n_st = L.vl_profile_stage_count()
      ms_arr = (ctypes.c_double * n_st)()
      cnt_arr = (ctypes.c_longlong * n_st)()
      L.vl_profile_collect(ms_arr, cnt_arr)
      stage = {L.vl_profile_stage_name(i).decode(): (ms_arr[i], cnt_arr[i]) for i in range(n_st) if cnt_arr[i]}
    if world > 1:
      t = torch.tensor([ms], dtype=torch.float64, device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    return ms, launches, stage

  import ctypes
  for _ in range(Wm):  # warm-up: both legs
    step_device()
  rr.wait()
  for _ in range(min(Wm, 3)):
    step_host()
  rr.wait()

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  ms_dev, launches, _ = timed(step_device, K, profile=False)
  ms_e2e, _, _ = timed(step_host, K, profile=False)
  clocks = sampler.stop() if rank == 0 else None
  # per-kernel durations: same steps again with the library's event profiler on (events are recorded on the
  # launching streams; kept out of the headline timing because each record costs host time per launch)
  # -- on ONE stream, so that a kernel's duration is not inflated by kernels of other scans sharing the SMs
  rr1 = pipeline.ScanRenderer(rays_np, origin, H, max_v, max_f, n_streams=1, device=dev, host_io=False,
                              method=args.method, use_graph=False)

  def step_profile():
    for ds in d_scenes:
      rr1.submit(*ds)
  ms_prof, _, stage = timed(step_profile, max(1, min(K, 4)), profile=True)
  n_active = n_units = 0.0
  if args.method == "cast":   # triangles that can be hit at all / work units of the last profiled scan
    info = (ctypes.c_int * 8)()
    L.vl_cast_status(ctypes.c_void_p(rr1.slots[0].blob.data_ptr()), ctypes.c_void_p(rr1.slots[0].stream.cuda_stream), info)
    n_active, n_units = float(info[1]), float(info[2])

  # the other device path on the same scans, briefly (same bits, different structure: tests/test_cast_gpu.py)
  other = "lbvh" if args.method == "cast" else "cast"
  rr2 = pipeline.ScanRenderer(rays_np, origin, H, max_v, max_f, n_streams=args.streams, device=dev, host_io=False,
                              method=other)

  def step_other():
    for k in range(S):
      rr2.submit(*d_scenes[k % M])
  for _ in range(3):
    step_other()
  rr2.wait()
  ms_other, _, _ = timed(step_other, max(1, min(K, 5)), profile=False, fence=rr2)
  other_value = S * H * W * world * max(1, min(K, 5)) / (ms_other * 1e-3) / 1e6

  # parity spot check of the last scan against the library's brute-force kernel on a ray subset is done in
  # tests/; here only a cheap sanity check that rays hit
  rr.wait()
  hit_frac = float((rr.slots[(S - 1) % len(rr.slots)].out["tri_id"] >= 0).float().mean().item())

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  rays_per_step = S * H * W * world
  value = rays_per_step * K / (ms_dev * 1e-3) / 1e6
  e2e_value = rays_per_step * K / (ms_e2e * 1e-3) / 1e6
  h2d = sum(mesh_bytes[k % M] for k in range(S))   # per rank per step
  d2h = S * (H * W) * (12 + 12 + 4 + 4 + 4)
  peak, peak_src = _peaks()

  # roofline of the dominant kernel (algorithmic bytes per launch, DESIGN.md "kernels")
  nt = float(np.mean(n_tris)); nv = float(np.mean(n_verts)); R = H * W
  alg_bytes = {
      "bounds": 12 * nv,
      "morton": 12 * nt + 12 * nv + 4 * nt + 4 * nt,          # faces + verts(gather, once) + key + flag
      "sort_pass": (4 + 4) * nt * 2,                          # key+val in, key+val out (one 8-bit digit)
      "emit_climb": 8 * nt + 12 * nt + 28 * nv + 48 * nt + 16 * nt + 64 * 0.3 * nt,   # ~0.3 nodes per triangle are written
      "top_climb": 64 * 0.004 * nt,
      "trace": 112 * nt + 12 * R + 36 * R,
      # scene-streaming cast: faces + vertices read once; records (64 B) and work units (8 B) of the triangles that
      # can be hit at all written once, read once; beam index (sorted beams 16 B, cells 4 B, slots 8 B) read once
      "cast_init": 8 * R,
      "cast_setup": 12 * nt + 12 * nv + 64 * n_active + 8 * n_units,
      "cast_items": 64 * n_active + 8 * n_units + 16 * R + 4 * R + 8 * R,
      "cast_resolve": (8 + 4 + 16) * R + 36 * R + 36 * R,   # keys, slots, directions in; outputs; face/colour/remission gathers
  }
  path = ("cast_init", "cast_setup", "cast_items", "cast_resolve") if args.method == "cast" else (
      "bounds", "morton", "sort_pass", "emit_climb", "top_climb", "trace")
  total_stage_ms = sum(v[0] for v in stage.values()) or 1.0
  dom = max(stage.items(), key=lambda kv: kv[1][0])[0]
  dom_ms, dom_n = stage[dom]
  achieved = alg_bytes.get(dom, 0.0) / (dom_ms / dom_n * 1e-3) / 1e9
  stages_out = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share": v[0] / total_stage_ms,
                    "alg_GBps": alg_bytes.get(k, 0.0) / (v[0] / v[1] * 1e-3) / 1e9} for k, v in stage.items()}

  pipe = None
  if world == 1 and not args.no_pipeline:
    del rr2
    torch.cuda.empty_cache()
    pipe = pipeline_c1(L, dev)

  # cpu baseline: the reference C++ ray tracer on a bounded sample of the same scans
  cpu = None
  if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
    n_calls = min(args.cpu_scans, M)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    times, kind = time_reference(scenes, rays_np, n_calls=n_calls, warmup=0)
    cpu_val = n_calls * H * W / sum(times) / 1e6
    cpu = {"value": cpu_val, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": kind,
           "sample": "%d of the step's %d distinct scans (%d tris, %d rays each), one ctrace call per scan incl. triangle "
                     "construction + BVH build; %.2f s/scan" % (n_calls, M, n_tris[0], R, sum(times) / n_calls)}

  line = {
      "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
      "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f32", "data": "synthetic",
      "config": {"workload": WORKLOAD, "method": args.method, "scans_per_step_per_gpu": S, "rays_per_scan": R,
                 "tris_per_scan": int(nt), "streams": args.streams,
                 "l2": "inputs larger than L2: %d distinct meshes x %.0f MB cycled per step + %.0f MB scratch per stream"
                       % (M, mesh_bytes[0] / 1e6, rr.slots[0].blob.numel() / 1e6),
                 "submission": "one CUDA graph launch per scan (vl_cast_graph_launch)" if rr.use_graph else "kernel by kernel",
                 "e2e_api": "ScanRenderer.submit_host (pinned host mesh -> H2D -> %s -> D2H of 5 outputs)"
                            % ("vl_cast" if args.method == "cast" else "vl_bvh_build -> vl_trace"),
                 "beam_index": "built once per sensor outside the timed region (vl_beams_build, ~40 us)" if args.method == "cast" else None},
      "scans_per_s": S * world * K / (ms_dev * 1e-3),
      "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
              "scans_per_s": S * world * K / (ms_e2e * 1e-3), "ms_per_step": ms_e2e / K},
      "gpu_launches": int(launches),
      "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                   "frac": achieved / peak, "traffic": _ncu_traffic(dom), "peak_source": peak_src,
                   "alg_bytes_per_launch": alg_bytes.get(dom, 0.0), "ms_per_launch": dom_ms / dom_n,
                   "step_alg_bytes": sum(alg_bytes[k] * (4 if k == "sort_pass" else 1) for k in path) * S,
                   "scan_alg_bytes_in_out": 12 * nt + 28 * nv + 36 * R,
                   "tris_that_can_be_hit": n_active, "work_units": n_units},
      "stages": stages_out,
      "other_method": {"method": other, "value": other_value, "unit": "Mrays/s", "ms_per_step": ms_other / max(1, min(K, 5))},
      "pipeline_c1": pipe,
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
  ap.add_argument("--scans-per-step", type=int, default=32)
  ap.add_argument("--distinct-meshes", type=int, default=8)
  ap.add_argument("--streams", type=int, default=8)
  ap.add_argument("--method", default="cast", choices=["cast", "lbvh"])
  ap.add_argument("--cpu-scans", type=int, default=8, help="scans timed for cpu_baseline (about 1.2 s each)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-pipeline", action="store_true", help="skip the config-1 chain measurement (pipeline_c1)")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
  if args.impl == "reference":
    run_reference(args)
  else:
    run_native(args)


if __name__ == "__main__":
  main()
