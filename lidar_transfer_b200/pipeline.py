"""One-scan-per-stream executor for batches of independent scans.

Scans are independent units (every scan has its own mesh): a ScanRenderer owns `n_streams` CUDA streams,
each with pre-allocated scratch, staging and output buffers, and round-robins scans over them -- no
allocation and no host synchronisation in steady state.  Two device paths give the same bits:
method="cast" (default) indexes the sensor's beams ONCE and streams every scan's triangles through that
index (vl_beams_build + vl_cast); method="lbvh" builds a per-scan LBVH and traverses it per ray
(vl_bvh_build + vl_trace).
The reference processes scans one by one in its driver loop (lidar_deform.py:393-458 ->
TSDFVolume.throw_rays_at_mesh, auxiliary/fusion_lidar.py:426-455 -> C_Trace); this is the batch
form of the same call, also used to shard scans across GPUs (sharding.py).
"""
import ctypes

import numpy as np
import torch

from . import engine
from . import _lib as _lib_mod
from ._lib import check, lib


def _ptr(t):
  return ctypes.c_void_p(t.data_ptr())


class _Slot:
  def __init__(self, dev, max_verts, max_faces, n_rays, host_io, method):
    f32, i32 = torch.float32, torch.int32
    self.stream = torch.cuda.Stream(device=dev)
    # per-scan scratch: the LBVH blob, or the cast workspace (8 B per beam)
    need = lib().vl_bvh_blob_bytes(max_faces) if method == "lbvh" else lib().vl_cast_workspace_bytes(n_rays, max_faces)
    self.blob = torch.empty(need, dtype=torch.uint8, device=dev)
    self.out = dict(endpoints=torch.empty(3 * n_rays, dtype=f32, device=dev),
                    endcolors=torch.empty(3 * n_rays, dtype=i32, device=dev),
                    range=torch.empty(n_rays, dtype=f32, device=dev),
                    endrem=torch.empty(n_rays, dtype=f32, device=dev),
                    tri_id=torch.empty(n_rays, dtype=i32, device=dev))
    self.done = torch.cuda.Event()
    self.busy = False
    self.inputs = None   # the mesh tensors of the scan in flight: held until it has drained (see ScanRenderer.submit)
    self.owner = None
    # bytes 16..31 of the cast workspace header (n_bad_faces, overflow, ... as k_cast_resolve left them), copied back after every scan
    self.h_status = torch.zeros(4, dtype=torch.int32, pin_memory=True)
    self.h_status_np = self.h_status.numpy()
    self.d_status = self.blob[16:32].view(torch.int32)
    # raw handles for the single-call submission path (vl_cast_submit): events must exist before their handle does
    self.ready = torch.cuda.Event()
    self.ready.record(self.stream)
    self.done.record(self.stream)
    self.fixed = (self.out["endpoints"].data_ptr(), self.out["endcolors"].data_ptr(), self.out["range"].data_ptr(),
                  self.out["endrem"].data_ptr(), self.out["tri_id"].data_ptr(), engine.TRACE_ZERO_MISSES,
                  self.blob.data_ptr(), self.blob.numel(), self.stream.cuda_stream)
    self.tail = (self.ready.cuda_event, self.h_status.data_ptr(), self.done.cuda_event)
    # the mesh descriptor of the graph path (vl_cast_graph_*): 4 device pointers, n_verts, n_faces
    self.h_desc = torch.zeros(8, dtype=torch.int64, pin_memory=True)
    self.desc64 = self.h_desc.numpy()
    self.desc32 = self.desc64.view(np.int32)
    self.graph = None
    if host_io:
      # device staging for host-fed meshes + pinned host buffers for the results
      self.d_verts = torch.empty(3 * max_verts, dtype=f32, device=dev)
      self.d_faces = torch.empty(3 * max_faces, dtype=i32, device=dev)
      self.d_colors = torch.empty(3 * max_verts, dtype=i32, device=dev)
      self.d_rem = torch.empty(max_verts, dtype=f32, device=dev)
      self.h_out = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in self.out.items()}

  def result(self):
    """Waits for THIS slot's scan, raises VlidarError(VL_ENOSPACE) if its cast ran out of work units (outputs invalid;
    the mesh needs ScanRenderer(method='lbvh')), releases the input tensors and returns the output dict (`h_out` for a
    host-fed scan).  Idempotent until the slot is reused."""
    if self.busy:
      self.done.synchronize()
      self.busy = False
      self.inputs = None
      self.owner._check(self)
    return getattr(self, "h_out", None) if self.owner.host_io and self.host_fed else self.out


class ScanRenderer:
  """rays f32[R,3] and origin f32[3] are fixed per renderer (one target sensor)."""

  def __init__(self, rays, origin, height, max_verts, max_faces, n_streams=4, device=None, host_io=False,
               method="cast", use_graph=True, normalize=None):
    engine.require_cuda()
    if method not in ("cast", "lbvh"):
      raise ValueError("method must be 'cast' or 'lbvh'")
    self.method = method
    self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    # directions normalised once per sensor (engine.DEFAULT_NORMALIZE: on the host, the reference's own arithmetic)
    self.rays, self.ray_flags = engine._prepare_rays(rays, normalize, self.dev)
    self.origin = engine._dev(origin, torch.float32, self.dev).reshape(-1)
    self.n_rays = self.rays.numel() // 3
    self.height = int(height)
    self.max_verts, self.max_faces = int(max_verts), int(max_faces)
    self.slots = [_Slot(self.dev, max_verts, max_faces, self.n_rays, host_io, method) for _ in range(n_streams)]
    for s in self.slots:
      s.owner, s.host_fed = self, False
    # the beam index depends on the sensor only: built once here, shared (read-only) by every stream
    self.beams = (engine.Beams(self.rays, self.height, self.dev, normalize="given" if self.ray_flags else "ieee")
                  if method == "cast" else None)
    if self.beams is not None:
      torch.cuda.current_stream(self.dev).synchronize()
    self.host_io = host_io
    self._dev_index = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
    self._next = 0
    self._lib = lib()
    self._beams_ptr = self.beams.blob.data_ptr() if self.beams is not None else 0
    self._origin_ptr = self.origin.data_ptr()
    # method "cast": every slot's four kernels + status copy are captured once and replayed per scan with a new mesh
    # descriptor (one graph launch instead of six driver calls).  Set use_graph = False to launch kernel by kernel
    # (the per-stage event profiler of the library needs that).
    self.use_graph = bool(use_graph) and method == "cast"
    if self.use_graph:
      with torch.cuda.device(self.dev):
        for s in self.slots:
          g = ctypes.c_void_p()
          check(self._lib.vl_cast_graph_create(self._beams_ptr, self._origin_ptr, self.n_rays, self.height, *s.fixed[:8],
                                               self.max_faces, s.h_desc.data_ptr(), s.h_status.data_ptr(),
                                               s.stream.cuda_stream, ctypes.byref(g)))
          s.graph = g

  def close(self):
    """Destroys the slots' graphs (after everything submitted has drained)."""
    self.wait()
    for s in self.slots:
      if s.graph is not None:
        self._lib.vl_cast_graph_destroy(s.graph)
        s.graph = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def _acquire(self):
    s = self.slots[self._next]
    self._next = (self._next + 1) % len(self.slots)
    if s.busy:
      s.result()  # the slot's previous scan must have drained before its buffers are reused (raises if it overflowed)
    return s

  def _launch(self, s, verts, faces, colors, rem, n_verts, n_faces):
    L = self._lib
    st = ctypes.c_void_p(s.stream.cuda_stream)
    if self.method == "cast":
      check(L.vl_cast(_ptr(self.beams.blob), _ptr(verts), _ptr(faces), _ptr(colors), _ptr(rem), n_verts, n_faces,
                      _ptr(self.origin), self.n_rays, self.height, _ptr(s.out["endpoints"]), _ptr(s.out["endcolors"]),
                      _ptr(s.out["range"]), _ptr(s.out["endrem"]), _ptr(s.out["tri_id"]), engine.TRACE_ZERO_MISSES,
                      _ptr(s.blob), s.blob.numel(), st))
      with torch.cuda.stream(s.stream):
        s.h_status.copy_(s.d_status, non_blocking=True)
      return
    check(L.vl_bvh_build(_ptr(verts), _ptr(faces), _ptr(colors), _ptr(rem), n_verts, n_faces, _ptr(s.blob),
                         s.blob.numel(), st))
    check(L.vl_trace(_ptr(s.blob), n_faces, _ptr(self.rays), _ptr(self.origin), self.n_rays, self.height,
                     _ptr(s.out["endpoints"]), _ptr(s.out["endcolors"]), _ptr(s.out["range"]), _ptr(s.out["endrem"]),
                     _ptr(s.out["tri_id"]), engine.TRACE_ZERO_MISSES | self.ray_flags, st))

  def submit(self, verts, faces, colors, rem):
    """Device-resident mesh (flat CUDA tensors: verts f32[3N_v], faces i32[3N_t], colors i32[3N_v],
    rem f32[N_v]).  Asynchronous; returns the slot: `slot.result()` waits for this scan, checks it and returns the
    output tensors (`slot.out`; valid until the slot is reused, n_streams submissions later).
    Lifetime: the kernels read the four tensors on the slot's own stream, which the caching allocator does not know
    about -- the slot therefore keeps references to them until the scan has drained (result() / wait() / reuse of the
    slot), so the caller may drop or overwrite its own references right after submit()."""
    n_verts, n_faces = verts.numel() // 3, faces.numel() // 3
    if n_faces > self.max_faces:
      raise ValueError("mesh has %d faces, renderer was sized for %d" % (n_faces, self.max_faces))
    s = self._acquire()
    s.inputs, s.host_fed = (verts, faces, colors, rem), False
    if torch.cuda.current_device() != self._dev_index:   # launches go to the renderer's device whatever the caller's is
      with torch.cuda.device(self.dev):
        return self._submit_on_device(s, verts, faces, colors, rem, n_verts, n_faces)
    return self._submit_on_device(s, verts, faces, colors, rem, n_verts, n_faces)

  def _submit_on_device(self, s, verts, faces, colors, rem, n_verts, n_faces):
    if self.use_graph and s.graph is not None:
      # one graph launch per scan: the mesh goes through the slot's pinned descriptor
      d64, d32 = s.desc64, s.desc32
      d64[0] = verts.data_ptr(); d64[1] = faces.data_ptr(); d64[2] = colors.data_ptr(); d64[3] = rem.data_ptr()
      d32[8] = n_verts; d32[9] = n_faces
      check(self._lib.vl_cast_graph_launch(s.graph, s.fixed[8], torch.cuda.current_stream(self.dev).cuda_stream,
                                           s.tail[0], s.tail[2]))
      s.busy = True
      return s
    if self.method == "cast":
      # one C call per scan: wait for the producer stream, four launches, status copy, done event
      # (the caller is on self.dev; per-scan host time bounds the batch at this kernel speed)
      check(self._lib.vl_cast_submit(self._beams_ptr, verts.data_ptr(), faces.data_ptr(), colors.data_ptr(), rem.data_ptr(),
                                     n_verts, n_faces, self._origin_ptr, self.n_rays, self.height, *s.fixed,
                                     torch.cuda.current_stream(self.dev).cuda_stream, *s.tail))
      s.busy = True
      return s
    s.stream.wait_stream(torch.cuda.current_stream(self.dev))
    self._launch(s, verts, faces, colors, rem, n_verts, n_faces)
    s.done.record(s.stream)
    s.busy = True
    return s

  def submit_host(self, verts, faces, colors, rem):
    """Host mesh (pinned or pageable torch CPU tensors / numpy arrays, flat, reference dtypes):
    H2D on the slot's stream, build, trace, D2H of the five outputs into the slot's pinned buffers.  The slot keeps
    references to the host tensors until the scan has drained (the copies are asynchronous for pinned memory)."""
    if not self.host_io:
      raise RuntimeError("renderer was created without host_io=True")
    as_t = lambda a: torch.from_numpy(a) if isinstance(a, np.ndarray) else a
    verts, faces, colors, rem = (as_t(a).reshape(-1) for a in (verts, faces, colors, rem))
    n_verts, n_faces = verts.numel() // 3, faces.numel() // 3
    if n_faces > self.max_faces or n_verts > self.max_verts:
      raise ValueError("mesh (%d verts, %d faces) exceeds the renderer's capacity" % (n_verts, n_faces))
    s = self._acquire()
    s.inputs, s.host_fed = (verts, faces, colors, rem), True
    with torch.cuda.device(self.dev), torch.cuda.stream(s.stream):
      s.d_verts[:3 * n_verts].copy_(verts, non_blocking=True)
      s.d_faces[:3 * n_faces].copy_(faces, non_blocking=True)
      s.d_colors[:3 * n_verts].copy_(colors, non_blocking=True)
      s.d_rem[:n_verts].copy_(rem, non_blocking=True)
      self._launch(s, s.d_verts, s.d_faces, s.d_colors, s.d_rem, n_verts, n_faces)
      for k, v in s.out.items():
        s.h_out[k].copy_(v, non_blocking=True)
    s.done.record(s.stream)
    s.busy = True
    return s

  def _check(self, s):
    """After the slot's scan has drained: the cast must not have run out of work units (its outputs would be all
    misses); such a mesh has to go through method="lbvh"."""
    if self.method == "cast" and s.h_status_np[1] != 0:
      raise _lib_mod.VlidarError(_lib_mod.VL_ENOSPACE, "the mesh needs more cast work units than the workspace holds; "
                                 "use ScanRenderer(method='lbvh') for it")

  def wait(self):
    err = None
    for s in self.slots:
      try:
        s.result()
      except _lib_mod.VlidarError as e:   # drain every slot before reporting the first failure
        err = err or e
    if err is not None:
      raise err

  def fence(self):
    """Make the current torch stream wait for everything submitted so far (device-side join)."""
    cur = torch.cuda.current_stream(self.dev)
    for s in self.slots:
      cur.wait_stream(s.stream)


class _Lane:
  """One scan in flight of a ScanPipeline: its own stream, TSDF volume, projection workspace and result buffers."""

  def __init__(self, owner):
    dev = owner.dev
    self.stream = torch.cuda.Stream(device=dev)
    with torch.cuda.device(dev):
      self.vol = engine.TsdfDevice(owner.dim, owner.vol_origin, owner.voxel_size, owner.fov_up, owner.fov_down)
    R = owner.n_rays
    self.packed = torch.empty(32 * R, dtype=torch.uint8, device=dev)   # endpoints | endcolors | range | endrem
    self.outs = dict(endpoints=self.packed[:12 * R].view(torch.float32), endcolors=self.packed[12 * R:24 * R].view(torch.int32),
                     range=self.packed[24 * R:28 * R].view(torch.float32), endrem=self.packed[28 * R:32 * R].view(torch.float32))
    self.h_out = torch.empty(32 * R, dtype=torch.uint8).pin_memory()
    self.done = torch.cuda.Event()
    self.ws = None
    self.proj = None     # the projection's output tensors, written again by every scan of this lane
    self.color_im = torch.empty((owner.im_h, owner.im_w), dtype=torch.float32, device=dev)
    self.mesh_buf = {}   # grow-only storage of the mesh arrays (engine.TsdfDevice.extract_mesh_finish)
    self.cast_ws = None  # grow-only cast workspace
    self.ctx = None      # between the halves: the mesh count is on its way
    self.tag = None
    self.mesh = None     # the mesh of the scan whose cast is in flight (kept until the next scan of this lane)
    self.inputs = None


class ScanPipeline:
  """The whole per-scan chain for batches of independent scans, points in -> per-ray results out:
  projection (vl_project) -> sparse TSDF (vl_tsdf_sparse_integrate) -> iso-surface (vl_mesh_*) -> cast (vl_cast),
  i.e. what MultiSemLaserScan.deform('mergemesh') does for one scan (auxiliary/laserscan.py:921-1012), software-pipelined
  over `n_lanes` scans in flight: the chain has ONE host synchronisation per scan (the triangle count that sizes the
  mesh), and while the host waits for the count of scan k the kernels of scan k + 1 are already queued on another
  stream.  Same kernels, same bits as the sequential engine calls (tests/test_chain_gpu.py).

      pipe = ScanPipeline(rays, height, src_fov, vol_bnds, voxel_size, im_h, im_w)
      for tag, h_out in pipe.run(clouds):      # clouds: iterable of (points f64[N,3] or f32[N,3], remission f32[N], label i32[N])
        ...                                    # h_out: pinned uint8[32 R] = endpoints | endcolors | range | endrem

  Scans are independent, so a multi-GPU job gives each rank its own ScanPipeline over its shard (sharding.py)."""

  def __init__(self, rays, height, fov_up, fov_down, vol_bnds, voxel_size, im_h, im_w, n_lanes=3, device=None, origin=None,
               expect_tris=0):
    """expect_tris: size every lane's mesh arrays and cast workspace for this many triangles up front (they grow by
    reallocation otherwise: the first ~100 scans of a process then run at half speed while the buffers find their size)."""
    engine.require_cuda()
    self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    self.fov_up, self.fov_down = float(fov_up), float(fov_down)
    self.im_h, self.im_w = int(im_h), int(im_w)
    bnds = np.asarray(vol_bnds, np.float64).reshape(3, 2)
    self.dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / voxel_size).astype(int)   # fusion_lidar.py:39-41
    self.vol_origin = bnds[:, 0].astype(np.float32)
    self.voxel_size = float(voxel_size)
    self.height = int(height)
    self.beams = engine.Beams(rays, self.height, self.dev)
    self.n_rays = self.beams.n_rays
    self.origin = torch.zeros(3, device=self.dev) if origin is None else engine._dev(origin, torch.float32, self.dev).reshape(-1)
    self.lanes = [_Lane(self) for _ in range(max(1, int(n_lanes)))]
    if expect_tris > 0:
      cap = int(expect_tris)
      for lane in self.lanes:
        with torch.cuda.stream(lane.stream):
          lane.mesh_buf.update(cap_t=cap, verts=torch.empty((3 * cap, 3), dtype=torch.float32, device=self.dev), faces=None, norms=None,
                               colors=torch.empty((3 * cap, 3), dtype=torch.uint8, device=self.dev),
                               rem=torch.empty(3 * cap, dtype=torch.float32, device=self.dev))
          lane.cast_ws = self.beams.workspace(cap)
    torch.cuda.synchronize(self.dev)

  def _front(self, lane, tag, cloud):
    """points -> range / label / remission image -> TSDF -> triangle count on its way to the host"""
    with torch.cuda.stream(lane.stream):
      p64, rem, lab = (t.to(self.dev, non_blocking=True) if torch.is_tensor(t) else t for t in cloud)
      if torch.is_tensor(p64) and p64.dtype == torch.float32:
        p64 = p64.to(torch.float64)   # a scan file's float32 coordinates travel as they are (half the bytes) and are widened here: exact
      lane.inputs = (p64, rem, lab)
      pr = lane.proj = engine.project(p64, rem, lab, self.fov_up, self.fov_down, self.im_h, self.im_w, workspace=lane.ws,
                                      out=lane.proj, want_keep=False)
      lane.ws = pr["workspace"]
      lane.color_im.copy_(pr["proj_label"])   # the folded single-channel colour image (label * 65536, fusion_lidar.py:263-264)
      lane.color_im.mul_(65536.0)
      lane.vol.reset()
      lane.vol.integrate(lane.color_im, pr["range_image"], pr["proj_remissions"])
      lane.ctx = lane.vol.extract_mesh_begin()
      lane.tag = tag

  def _back(self, lane):
    """triangle count -> mesh -> cast -> results on their way to pinned host memory"""
    with torch.cuda.stream(lane.stream):
      m = lane.vol.extract_mesh_finish(lane.ctx, want_norms=False, buffers=lane.mesh_buf, want_faces=False)   # a soup: no index array
      n_faces = m["n_tris"]
      if lane.cast_ws is None or lane.cast_ws.numel() < lib().vl_cast_workspace_bytes(self.n_rays, n_faces):
        lane.cast_ws = self.beams.workspace(n_faces + n_faces // 4)
      engine.cast(self.beams, m["verts"], None, m["colors"], m["rem"], self.origin, out=lane.outs, want_ids=False,
                  zero_misses=True, check_mesh=False, workspace=lane.cast_ws)
      lane.h_out.copy_(lane.packed, non_blocking=True)
      lane.done.record(lane.stream)
      lane.mesh, lane.ctx = m, None

  def run(self, clouds):
    """Generator over (tag, pinned result buffer) in submission order; `clouds` yields (points, remission, label) or
    (tag, points, remission, label).  A yielded buffer is valid until its lane is reused, n_lanes scans later."""
    n = len(self.lanes)
    for lane in self.lanes:   # a previous run that was abandoned half way (an exception in the consumer) leaves no trace
      lane.ctx = None
    waiting = []   # lanes in flight, oldest first (round robin: a lane that is needed again is always the oldest)
    k = 0

    def hand_out(w):
      if w.ctx is not None:
        self._back(w)
      w.done.synchronize()
      return w.tag, w.h_out

    for item in clouds:
      tag, cloud = (item[0], item[1:]) if len(item) == 4 else (k, item)
      lane = self.lanes[k % n]
      if waiting and waiting[0] is lane:
        yield hand_out(waiting.pop(0))
      self._front(lane, tag, cloud)
      # every earlier scan still waiting for its count gets its second half now: the count has had a whole front of
      # another scan to arrive, and the kernels just queued keep the device busy while the host allocates and launches
      for w in waiting:
        if w.ctx is not None:
          self._back(w)
      waiting.append(lane)
      k += 1
    for w in waiting:
      yield hand_out(w)
