"""GPU parity of row (vi) / N3: identity re-render metrics on the device (vl_compare) vs the oracle's restatement of
compare() + iouEval, which tests/test_oracle_pinned.py pins to the reference's own functions."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
KEYS = ("source_color", "target_color", "source_label", "target_label", "source_range", "target_range", "source_rem", "target_rem")


def _check(engine, oracle, a, nclasses):
  ref = oracle.compare_numpy(nclasses=nclasses, **a)
  got = engine.compare(*[a[k] for k in KEYS], nclasses)
  assert np.array_equal(got["conf"], ref["conf"])
  assert got["n_present"] == ref["n_present"]
  assert got["m_iou"] == ref["m_iou"] and got["m_acc"] == ref["m_acc"]                       # same matrix, same formulas
  assert abs(got["mse"] - float(ref["mse"])) <= 1e-6 * max(1.0, float(ref["mse"]))               # double vs pairwise float32 sum
  assert np.array_equal(got["range_diff"].cpu().numpy().view(np.int32), ref["range_diff"].view(np.int32))
  assert np.array_equal(got["rem_diff"].cpu().numpy().view(np.int32), ref["rem_diff"].view(np.int32))
  assert np.abs(got["label_diff"].cpu().numpy() - ref["label_diff"]).max() <= 1e-6             # float32 vs float64 colours
  return got, ref


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_compare_matches_reference_golden_inputs(engine, oracle, tag):
  G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_compare_v1.npz"))
  a = {k: G["cmp_%s_%s" % (tag, k)] for k in KEYS}
  got, _ = _check(engine, oracle, a, int(G["cmp_%s_nclasses" % tag]))
  m_iou, m_acc, mse = G["cmp_%s_scalars" % tag]
  assert got["m_iou"] == m_iou and got["m_acc"] == m_acc and abs(got["mse"] - mse) <= 1e-6 * mse


def test_compare_full_image_identity_and_edge_cases(engine, oracle):
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
  from make_golden_compare import make_case
  a = make_case(9, 64, 2048, 15)
  _check(engine, oracle, a, 20)
  # identity: the target equals the source -> IoU = Acc = 1, MSE = 0
  ident = dict(a)
  for k in ("color", "label", "range", "rem"):
    ident["target_" + k] = a["source_" + k].copy()
  got, _ = _check(engine, oracle, ident, 20)
  assert got["m_iou"] == 1.0 and got["m_acc"] == 1.0 and got["mse"] == 0.0
  # everything black: one label (0) remains
  black = dict(a)
  black["source_color"] = np.zeros_like(a["source_color"])
  got, _ = _check(engine, oracle, black, 20)
  assert got["n_present"] == 1 and got["conf"][0, 0] == 64 * 2048


def test_compare_rejects_what_the_reference_cannot_index(engine):
  from lidar_transfer_b200._lib import VlidarError
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
  from make_golden_compare import make_case
  a = make_case(4, 16, 64, 15)
  with pytest.raises(VlidarError):      # 15 distinct labels, 5 classes: np.add.at raises IndexError in the reference
    engine.compare(*[a[k] for k in KEYS], 5)
  a["target_label"][3, 3] = 70000
  a["source_color"][3, 3] = 0.5
  a["source_label"][3, 3] = 10
  with pytest.raises(VlidarError):
    engine.compare(*[a[k] for k in KEYS], 20)
