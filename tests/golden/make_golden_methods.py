#!/usr/bin/env python
"""Generates tests/golden/golden_methods_v1.npz from the REFERENCE ITSELF: do_range_projection_new with the two methods
no caller of the reference selects (deform() always takes the default 'depth', auxiliary/laserscan.py:952):
  'pdist'     (:392-416)  per pixel the point whose image position is nearest to the pixel CENTRE;
  'depthfast' (:418-437)  per pixel the nearest point through argsort + fancy-index assignment (last write wins).
Same recipe as make_golden.py: the reference's own Python imported unmodified from /root/reference (stubs for the
absent imageio / skimage / matplotlib), run in the authoring container, result committed.

    python tests/golden/make_golden_methods.py

The scan is minimal.zip scan 0 decimated by 4 into a 16 x 128 image, so that most pixels see several points."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, load_minimal  # noqa: E402


def main():
  LS, _ = import_reference()
  scan, label, poses, cfg = load_minimal()
  dec = slice(0, None, 4)
  pts_in, rem_in, lab_in = scan[dec, :3].copy(), scan[dec, 3].copy(), (label[dec] & 0xFFFF).astype(np.uint32)
  fu, fd, H, W = 3.0, -25.0, 16, 128
  G = dict(args=np.array([fu, fd, H, W], np.float64), points_f32=pts_in, rem=rem_in, label=lab_in)
  for method in ("pdist", "depthfast"):
    for remove in (True, False):
      s = LS.SemLaserScan(H, W, 20, color_dict=cfg["color_map"])
      s.points, s.remissions, s.label = pts_in.astype(np.float64), rem_in.copy(), lab_in.copy()
      s.colorize()
      s.do_range_projection_new(fu, fd, remove=remove, method=method)
      k = "%s_%d_" % (method, int(remove))
      G[k + "n_kept"] = np.array([s.points.shape[0]], np.int64)
      if method == "pdist":
        G[k + "range"], G[k + "index"] = s.range_image.copy(), s.index.copy()
        G[k + "dist"] = s.dist_image.copy()
        G[k + "label"] = s.label_image[..., 0].copy()
        G[k + "rem"] = s.proj_remissions.copy()          # pdist never writes it: all -1
        G[k + "proj_y_float"] = np.asarray(s.proj_y_float).copy()
        print(k, "kept", s.points.shape[0], "pixels", int((s.index >= 0).sum()))
      else:
        G[k + "range"], G[k + "index"] = s.proj_range.copy(), s.proj_idx.copy()
        G[k + "rem"], G[k + "xyz"] = s.proj_remissions.copy(), s.proj_xyz.copy()
        # pixels whose two nearest points have the same depth: the winner there is decided by numpy's unstable argsort
        d = np.linalg.norm(s.points, 2, axis=1)
        print(k, "kept", s.points.shape[0], "pixels", int((s.proj_idx >= 0).sum()), "distinct depths", np.unique(d).size, "of", d.size)
  out = os.path.join(HERE, "golden_methods_v1.npz")
  np.savez_compressed(out, **G)
  print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
  main()
