"""Host I/O beside the device path: prefetching reader and asynchronous writer.  CPU only."""
import os
import struct

import numpy as np
import pytest

from lidar_transfer_b200 import scanio


def _make_files(tmp_path, n_frames, rng, with_labels=True):
  scans, labels, data = [], [], []
  for k in range(n_frames):
    n = int(rng.integers(50, 4000))
    pts = rng.normal(0, 20, (n, 4)).astype(np.float32)
    lab = (rng.integers(0, 260, n).astype(np.uint32) | (rng.integers(0, 9, n).astype(np.uint32) << 16))
    p, l = str(tmp_path / ("%06d.bin" % k)), str(tmp_path / ("%06d.label" % k))
    pts.tofile(p); lab.tofile(l)
    scans.append(p); labels.append(l); data.append((pts, lab))
  return scans, (labels if with_labels else None), data


@pytest.mark.parametrize("depth", [1, 3])
def test_prefetcher_yields_every_frame_in_order(tmp_path, depth):
  rng = np.random.default_rng(0)
  scans, labels, data = _make_files(tmp_path, 12, rng)
  prev = None
  with scanio.ScanPrefetcher(scans, labels, depth=depth, initial_points=100) as pf:   # forces the ring to grow
    seen = []
    for idx, scan, lab in pf:
      assert np.array_equal(scan, data[idx][0]) and np.array_equal(lab, data[idx][1])
      assert np.array_equal(scan, scanio.read_scan(scans[idx])) and np.array_equal(lab, scanio.read_label(labels[idx]))
      if prev is not None:   # the previous frame's views are still intact (its H2D copy may be in flight)
        assert np.array_equal(prev[1], data[prev[0]][0])
      prev = (idx, scan)
      seen.append(idx)
  assert seen == list(range(12))


def test_prefetcher_shards_like_scans_for_rank(tmp_path):
  from lidar_transfer_b200.sharding import scans_for_rank
  rng = np.random.default_rng(1)
  scans, _, data = _make_files(tmp_path, 11, rng, with_labels=False)
  got = []
  for rank in range(3):
    with scanio.ScanPrefetcher(scans, None, depth=2, start=rank, step=3) as pf:
      mine = [idx for idx, scan, lab in pf if lab is None and np.array_equal(scan, data[idx][0])]
    assert mine == list(scans_for_rank(11, rank, 3))
    got += mine
  assert sorted(got) == list(range(11))


def test_prefetcher_raises_reader_errors_at_their_frame(tmp_path):
  rng = np.random.default_rng(2)
  scans, labels, data = _make_files(tmp_path, 4, rng)
  data[2][1][:-1].tofile(labels[2])          # one label short
  os.remove(scans[3])
  with scanio.ScanPrefetcher(scans, labels, depth=2) as pf:
    assert next(pf)[0] == 0 and next(pf)[0] == 1
    with pytest.raises(ValueError, match="same number of points"):
      next(pf)
    with pytest.raises(OSError):
      next(pf)
    with pytest.raises(StopIteration):
      next(pf)
  with pytest.raises(ValueError):
    scanio.ScanPrefetcher(scans, labels[:2])


def test_async_writer_writes_the_reference_bytes(tmp_path):
  """The same bytes as the reference's per-point struct.pack loop (laserscan.py:1160-1178) after its filter rules
  (:1142-1158), for frames submitted faster than they are written."""
  rng = np.random.default_rng(3)
  out = str(tmp_path / "out")
  frames = []
  with scanio.AsyncScanWriter(out, max_pending=2, n_threads=2) as w:
    for k in range(9):
      n = 64 * 32
      pts = rng.normal(0, 10, (n, 3)).astype(np.float32)
      pts[rng.random(n) < 0.2] = 0.0                      # rays that hit nothing
      rem = rng.random(n).astype(np.float32)
      lab = rng.integers(-1, 260, n).astype(np.int64)     # -1: filtered out
      w.submit(k, pts.reshape(64, 32, 3), rem.reshape(64, 32), lab.reshape(64, 32))
      frames.append((pts, rem, lab))
  assert w.frames_written == 9
  for k, (pts, rem, lab) in enumerate(frames):
    want_bin, want_label = b"", b""
    for i in range(pts.shape[0]):                          # the reference's loop
      if lab[i] >= 0 and float(np.sum(pts[i])) != 0:
        want_bin += struct.pack("ffff", pts[i, 0], pts[i, 1], pts[i, 2], rem[i])
        want_label += struct.pack("I", int(lab[i]))
    assert open(os.path.join(out, "velodyne", "%06d.bin" % k), "rb").read() == want_bin
    assert open(os.path.join(out, "labels", "%06d.label" % k), "rb").read() == want_label


def test_async_writer_reports_worker_errors(tmp_path):
  w = scanio.AsyncScanWriter(str(tmp_path / "o"), max_pending=2)
  w.submit(0, np.zeros((4, 3), np.float32), np.zeros(4, np.float32), np.zeros(5, np.int32))   # shapes disagree
  with pytest.raises(Exception):
    w.close()
