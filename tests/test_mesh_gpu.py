"""GPU parity of the iso-surface extraction (vl_mesh_count / vl_mesh_emit) vs the oracle, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _load(dev_vol, name, arr):
  getattr(dev_vol, name).copy_(torch.from_numpy(arr))


@pytest.mark.parametrize("shape,seed", [((32, 30, 33), 0), ((70, 9, 130), 1), ((5, 4, 3), 2), ((2, 2, 2), 3)])
def test_mesh_extract_bit_exact(engine, oracle, shape, seed):
  rng = np.random.default_rng(seed)
  g = np.stack(np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij"), -1)
  c = np.array(shape) / 2.0 + 0.3
  vol = ((np.linalg.norm(g - c, axis=-1) - min(shape) * 0.37) / 3.0 + rng.normal(0, 0.05, shape)).astype(np.float32)
  vol = np.clip(vol, -1, 1)
  color_vol = (rng.choice([40, 48, 70, 259], size=shape) * 65536).astype(np.float32)
  rem_vol = rng.random(shape).astype(np.float32)
  vox, origin = 0.1, np.array([-1.5, 2.25, -3.0], np.float32)
  ref = oracle.mesh_extract(vol, color_vol, rem_vol, np.float32(vox), origin)
  dev = engine.TsdfDevice(shape, origin, vox, 3.0, -25.0)
  _load(dev, "tsdf", vol); _load(dev, "color", color_vol); _load(dev, "rem", rem_vol)
  got = dev.extract_mesh()
  assert got["faces"].shape[0] == ref["faces"].shape[0]
  if shape != (2, 2, 2):
    assert ref["faces"].shape[0] > 0
  assert np.array_equal(got["faces"].cpu().numpy(), ref["faces"])
  assert np.array_equal(got["verts"].cpu().numpy().view(np.int32), ref["verts"].view(np.int32))
  assert np.array_equal(got["colors"].cpu().numpy(), ref["colors"])
  assert np.array_equal(got["rem"].cpu().numpy().view(np.int32), ref["rem"].view(np.int32))
  n = got["norms"].cpu().numpy()
  ln = np.linalg.norm(n, axis=1)
  assert ((np.abs(ln - 1) < 1e-4) | (ln == 0)).all()


def test_empty_volume_gives_empty_mesh(engine):
  dev = engine.TsdfDevice((16, 16, 8), np.zeros(3, np.float32), 0.5, 3.0, -25.0)  # tsdf == 1 everywhere
  m = dev.extract_mesh()
  assert m["faces"].shape == (0, 3) and m["verts"].shape == (0, 3)
