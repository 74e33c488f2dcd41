#!/usr/bin/env python
"""Generates tests/golden/golden_compare_v1.npz from the REFERENCE ITSELF: compare()
(auxiliary/laserscan.py:1181-1301) with its own iouEval (auxiliary/np_ioueval.py), imported unmodified
from /root/reference through make_golden.import_reference(), on seeded synthetic images.

    python tests/golden/make_golden_compare.py

Inputs per case: a source scan (proj_color f64[H,W,3], proj_label i32[H,W], proj_range / proj_remissions f32[H,W])
and a re-rendered target scan (proj_color, label_image, proj_range, proj_remissions), with no-data pixels, source
background, label disagreements and labels missing on one side."""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

LABELS = np.array([0, 10, 11, 30, 40, 44, 48, 50, 51, 70, 71, 72, 80, 99, 252], np.int32)


def make_case(seed, H, W, n_labels):
  rng = np.random.default_rng(seed)
  labs = LABELS[:n_labels]
  lut = rng.random((256, 3))            # colour of a label (any positive floats)
  lut[0] = 0.0
  sl = rng.choice(labs, (H, W)).astype(np.int32)
  tl = np.where(rng.random((H, W)) < 0.8, sl, rng.choice(labs, (H, W))).astype(np.int32)
  if n_labels > 3:
    tl[tl == labs[2]] = labs[1]          # a label that occurs in the source only
  sc = lut[sl].copy()
  nodata = rng.random((H, W)) < 0.1      # black source pixels that still carry a label
  sc[nodata] = 0.0
  tc = lut[tl].copy()
  sr = np.where(sl > 0, rng.uniform(2, 80, (H, W)), 0).astype(np.float32)
  tr = (sr * rng.normal(1.0, 0.02, (H, W))).astype(np.float32)
  tr[rng.random((H, W)) < 0.05] = 0
  srem = rng.random((H, W)).astype(np.float32)
  trem = rng.random((H, W)).astype(np.float32)
  return dict(source_color=sc, target_color=tc, source_label=sl, target_label=tl, source_range=sr, target_range=tr,
              source_rem=srem, target_rem=trem)


def main():
  LS, FL = MG.import_reference()
  G = {}
  for tag, (seed, H, W, n_labels, nclasses) in dict(a=(1, 16, 64, 6, 20), b=(2, 64, 256, 15, 20), c=(3, 8, 32, 2, 5)).items():
    c = make_case(seed, H, W, n_labels)
    src = types.SimpleNamespace(proj_color=c["source_color"].copy(), proj_label=c["source_label"].copy(),
                                proj_range=c["source_range"].copy(), proj_remissions=c["source_rem"].copy(), nclasses=nclasses)
    tgt = types.SimpleNamespace(adaption="mesh", proj_color=c["target_color"].copy(), label_image=c["target_label"].copy(),
                                proj_range=c["target_range"].copy(), proj_remissions=c["target_rem"].copy())
    with contextlib.redirect_stdout(io.StringIO()):
      label_diff, range_diff, rem_diff, m_iou, m_acc, mse = LS.compare(src, tgt)
    for k, v in c.items():
      G["cmp_%s_%s" % (tag, k)] = v
    G["cmp_%s_nclasses" % tag] = np.array(nclasses)
    G["cmp_%s_label_diff" % tag] = label_diff
    G["cmp_%s_range_diff" % tag] = range_diff
    G["cmp_%s_rem_diff" % tag] = rem_diff
    G["cmp_%s_scalars" % tag] = np.array([m_iou, m_acc, mse], np.float64)
    print(tag, "m_iou %.6f m_acc %.6f mse %.6f" % (m_iou, m_acc, mse), label_diff.dtype, range_diff.dtype)
  out = os.path.join(HERE, "golden_compare_v1.npz")
  np.savez_compressed(out, **G)
  print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
  main()
