#!/usr/bin/env python
"""Generates tests/golden/golden_beams_v1.npz from the REFERENCE ITSELF: do_range_projection_new('depth') with the
`beam_angles` pitch snapping switched on (auxiliary/laserscan.py:321-327), which no fixture yaml of the reference
enables.  Same recipe as make_golden.py: the reference's own Python imported unmodified from /root/reference
(stubs for the absent imageio / skimage / matplotlib), run in the authoring container, result committed.

    python tests/golden/make_golden_beams.py

Two lists, both passed the way lidar_deform.py passes them (sorted python floats, laserscan.py:744):
  rad -- 16 angles in RADIANS across the field of view: what the comparison at :325 (pitch in radians minus the
         list) can meaningfully snap to;
  deg -- the same angles in DEGREES, which is what a yaml following config/lidar_transfer.yaml's comments would hold
         (the constructor's own note, laserscan.py:25: "TODO import deg transform to rad"): every pitch snaps to
         one of the two entries nearest zero (here -1.1 and 0.7), whose image rows lie outside [0, 1]: with remove=True no point survives
         and the reference stops with an IndexError at :384 (empty fancy index), so only remove=False is recorded.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, load_minimal  # noqa: E402


def main():
  LS, _ = import_reference()
  scan, label, poses, cfg = load_minimal()
  dec = slice(0, None, 16)
  pts_in, rem_in, lab_in = scan[dec, :3].copy(), scan[dec, 3].copy(), (label[dec] & 0xFFFF).astype(np.uint32)
  fu, fd, H, W = 3.0, -25.0, 16, 256
  deg = sorted(np.linspace(fd + 0.5, fu - 0.5, 16).tolist())
  rad = sorted((np.asarray(deg) / 180.0 * np.pi).tolist())
  G = dict(args=np.array([fu, fd, H, W], np.float64), points_f32=pts_in, rem=rem_in, label=lab_in,
           beams_rad=np.asarray(rad), beams_deg=np.asarray(deg))
  for tag, ba in (("rad", rad), ("deg", deg)):
    for remove in ((True, False) if tag == "rad" else (False,)):
      s = LS.SemLaserScan(H, W, 20, color_dict=cfg["color_map"], beam_angles=ba)
      s.points, s.remissions, s.label = pts_in.astype(np.float64), rem_in.copy(), lab_in.copy()
      s.colorize()
      s.do_range_projection_new(fu, fd, remove=remove)
      s.do_label_projection_new()
      k = "%s_%d_" % (tag, int(remove))
      G[k + "n_kept"] = np.array([s.points.shape[0]], np.int64)
      G[k + "range"], G[k + "index"] = s.range_image.copy(), s.index.copy()
      G[k + "label"], G[k + "rem"] = s.proj_label.copy(), s.proj_remissions.copy()
      G[k + "proj_y_float"] = np.asarray(s.proj_y_float).copy()
      print(k, "kept", s.points.shape[0], "pixels", int((s.index >= 0).sum()), "rows",
            np.unique(np.nonzero(s.index >= 0)[0]).size)
  out = os.path.join(HERE, "golden_beams_v1.npz")
  np.savez_compressed(out, **G)
  print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
  main()
