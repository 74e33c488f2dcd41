"""Drop-in for the reference's `auxiliary.np_ioueval.iouEval` (auxiliary/np_ioueval.py:8-70): confusion-matrix
IoU / accuracy used by compare() for the identity re-render self-check (rows = prediction, cols = target)."""
import numpy as np


class iouEval:
  def __init__(self, n_classes, ignore=None):
    self.n_classes = n_classes
    self.ignore = np.array(ignore, dtype=np.int64)
    self.include = np.array([n for n in range(self.n_classes) if n not in self.ignore], dtype=np.int64)
    print("[IOU EVAL] IGNORE: ", self.ignore)
    print("[IOU EVAL] INCLUDE: ", self.include)
    self.reset()

  def num_classes(self):
    return self.n_classes

  def reset(self):
    self.conf_matrix = np.zeros((self.n_classes, self.n_classes), dtype=np.int64)

  def addBatch(self, x, y):  # x=preds, y=targets
    x_row, y_row = np.asarray(x).reshape(-1), np.asarray(y).reshape(-1)
    assert x_row.shape == y_row.shape
    flat = x_row.astype(np.int64) * self.n_classes + y_row.astype(np.int64)
    self.conf_matrix += np.bincount(flat, minlength=self.n_classes ** 2).reshape(self.n_classes, self.n_classes)

  def getStats(self):
    conf = self.conf_matrix.copy()
    conf[self.ignore] = 0
    conf[:, self.ignore] = 0
    tp = np.diag(conf)
    return tp, conf.sum(axis=1) - tp, conf.sum(axis=0) - tp

  def getIoU(self):
    tp, fp, fn = self.getStats()
    union = tp + fp + fn + 1e-15
    iou = tp / union
    return (tp[self.include] / union[self.include]).mean(), iou

  def getacc(self):
    tp, fp, fn = self.getStats()
    return tp.sum() / (tp[self.include].sum() + fp[self.include].sum() + 1e-15)
