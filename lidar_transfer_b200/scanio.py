"""Host I/O beside the device path: the scan / label files of the next frames are read on a
background thread into a small ring of (optionally pinned) buffers while the device works on the current one, and
re-rendered scans are written out on another thread -- the reference reads with np.fromfile at the moment of use
(auxiliary/laserscan.py:116-140, 570-590) and writes point by point with struct.pack (:1160-1178), both on the one
Python thread that also drives the ray tracer.

On-disk formats are the reference's: KITTI `.bin` = float32[N,4] (x, y, z, remission), `.label` = uint32[N]
(lower 16 bits semantic class, laserscan.py:588).  Nothing here touches the GPU; pinned buffers only make the
subsequent host->device copy asynchronous."""
import os
import queue
import threading

import numpy as np


def read_scan(filename):
  """float32[N,4] of a KITTI .bin (laserscan.py:131-137)."""
  return np.fromfile(filename, dtype=np.float32).reshape((-1, 4))


def read_label(filename):
  """uint32[N] of a .label file, upper half (instance id) still in place (laserscan.py:583-588)."""
  return np.fromfile(filename, dtype=np.uint32).reshape((-1))


def filter_scan_for_write(back_points, remissions, label_image):
  """The filter rules of MultiSemLaserScan.write (laserscan.py:1142-1158): keep label >= 0, drop points whose
  x + y + z == 0 (rays that hit nothing).  Returns (float32[M,4] records, uint32[M] labels)."""
  back_points = np.asarray(back_points).reshape(-1, 3)
  remissions = np.asarray(remissions).reshape(-1)
  label_image = np.asarray(label_image).reshape(-1)
  valid = label_image >= 0
  back_points, remissions, label_image = back_points[valid], remissions[valid], label_image[valid].astype(np.int32)
  keep = np.sum(back_points, axis=1) != 0
  back_points, remissions, label_image = back_points[keep], remissions[keep], label_image[keep]
  rec = np.empty((back_points.shape[0], 4), dtype="<f4")
  rec[:, 0:3] = back_points
  rec[:, 3] = remissions
  return rec, label_image.astype("<u4")


def write_scan(out_dir, idx, rec, labels):
  """`velodyne/%06d.bin` + `labels/%06d.label`, the bytes struct.pack("ffff") / ("I") would produce (:1160-1178)."""
  rec.tofile(os.path.join(out_dir, "velodyne", str(idx).zfill(6) + ".bin"))
  labels.tofile(os.path.join(out_dir, "labels", str(idx).zfill(6) + ".label"))


class _Buffer:
  def __init__(self, n_points, pinned):
    self.capacity = 0
    self.pinned = pinned
    self._alloc(n_points)

  def _alloc(self, n_points):
    if self.pinned:
      import torch
      self._scan_t = torch.empty((n_points, 4), dtype=torch.float32).pin_memory()
      self._label_t = torch.empty((n_points,), dtype=torch.int32).pin_memory()
      self.scan, self.label = self._scan_t.numpy(), self._label_t.numpy().view(np.uint32)
    else:
      self.scan, self.label = np.empty((n_points, 4), np.float32), np.empty((n_points,), np.uint32)
    self.capacity = n_points


class ScanPrefetcher:
  """Iterates (idx, scan float32[N,4], label uint32[N] or None) over the given files IN ORDER; a background thread
  stays up to `depth` frames ahead.  The arrays of frame k are views into a ring buffer and remain valid until
  frame k + 1 has been requested twice over (i.e. for the current and the previous iteration step), so a consumer may
  still have the previous frame's host->device copy in flight.  pinned=True allocates the ring in page-locked memory
  (requires CUDA); errors in the reader thread (missing file, size mismatch) are raised at the frame they belong to."""

  def __init__(self, scan_names, label_names=None, depth=4, pinned=False, start=0, stop=None, step=1,
               initial_points=140000):
    if label_names is not None and len(label_names) != len(scan_names):
      raise ValueError("scan and label lists differ in length")
    self.scan_names, self.label_names = list(scan_names), (list(label_names) if label_names is not None else None)
    self.indices = list(range(len(self.scan_names)))[start:stop:step]   # e.g. start=rank, step=world for sharding
    self.depth = max(1, int(depth))
    self._free = queue.Queue()
    for _ in range(self.depth + 2):   # depth ahead + the current and the previous frame held by the consumer
      self._free.put(_Buffer(int(initial_points), pinned))
    self._ready = queue.Queue(maxsize=self.depth)
    self._held = []
    self._stop = threading.Event()
    self._thread = threading.Thread(target=self._run, name="scan-prefetch", daemon=True)
    self._thread.start()

  def _run(self):
    for idx in self.indices:
      buf = None
      while buf is None:
        if self._stop.is_set():
          return
        try:
          buf = self._free.get(timeout=0.1)
        except queue.Empty:
          continue
      try:
        n = os.path.getsize(self.scan_names[idx]) // 16
        if n > buf.capacity:
          buf._alloc(n + n // 8)
        with open(self.scan_names[idx], "rb") as f:
          got = f.readinto(memoryview(buf.scan[:n]).cast("B"))
        if got != 16 * n or os.path.getsize(self.scan_names[idx]) != 16 * n:
          raise ValueError("%s is not a float32[N,4] file" % self.scan_names[idx])
        has_label = self.label_names is not None
        if has_label:
          if os.path.getsize(self.label_names[idx]) != 4 * n:
            raise ValueError("Scan and Label don't contain same number of points")   # laserscan.py:586
          with open(self.label_names[idx], "rb") as f:
            f.readinto(memoryview(buf.label[:n]).cast("B"))
        item = (idx, buf, n, has_label, None)
      except Exception as e:   # handed to the consumer at this frame's position
        item = (idx, buf, 0, False, e)
      while not self._stop.is_set():
        try:
          self._ready.put(item, timeout=0.1)
          break
        except queue.Full:
          continue
    while not self._stop.is_set():
      try:
        self._ready.put(None, timeout=0.1)
        break
      except queue.Full:
        continue

  def __iter__(self):
    return self

  def __next__(self):
    item = self._ready.get()
    if item is None:
      self._ready.put(None)
      raise StopIteration
    idx, buf, n, has_label, err = item
    self._held.append(buf)
    while len(self._held) > 2:
      self._free.put(self._held.pop(0))
    if err is not None:
      raise err
    return idx, buf.scan[:n], (buf.label[:n] if has_label else None)

  def close(self):
    self._stop.set()
    self._thread.join(timeout=5.0)

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()


class AsyncScanWriter:
  """write() off the critical path: submit(idx, back_points, remissions, label_image) copies nothing (the caller
  hands the arrays over), a worker thread applies the reference's filter rules and writes the two files.  At most
  `max_pending` frames are queued (submit blocks beyond that); close() waits for all of them and re-raises the first
  error of the worker."""

  def __init__(self, out_dir, max_pending=8, n_threads=1):
    self.out_dir = out_dir
    os.makedirs(os.path.join(out_dir, "velodyne"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "labels"), exist_ok=True)
    self._q = queue.Queue(maxsize=max(1, int(max_pending)))
    self._err = None
    self.frames_written = 0
    self._lock = threading.Lock()
    self._threads = [threading.Thread(target=self._run, name="scan-writer-%d" % i, daemon=True) for i in range(max(1, n_threads))]
    for t in self._threads:
      t.start()

  def _run(self):
    while True:
      job = self._q.get()
      try:
        if job is None:
          return
        idx, pts, rem, lab = job
        rec, labels = filter_scan_for_write(pts, rem, lab)
        write_scan(self.out_dir, idx, rec, labels)
        with self._lock:
          self.frames_written += 1
      except Exception as e:
        with self._lock:
          if self._err is None:
            self._err = e
      finally:
        self._q.task_done()

  def submit(self, idx, back_points, remissions, label_image):
    if self._err is not None:
      raise self._err
    self._q.put((idx, back_points, remissions, label_image))

  def close(self):
    self._q.join()
    for _ in self._threads:
      self._q.put(None)
    for t in self._threads:
      t.join(timeout=10.0)
    if self._err is not None:
      raise self._err

  def __enter__(self):
    return self

  def __exit__(self, exc_type, *exc):
    if exc_type is None:
      self.close()
    else:   # do not mask the caller's exception
      try:
        self.close()
      except Exception:
        pass
