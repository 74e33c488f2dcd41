"""GPU parity of rows (i)+(ii): LBVH build + traversal vs the oracle, through the C ABI."""
import numpy as np
import pytest

from lidar_transfer_b200 import synth

pytestmark = pytest.mark.gpu

KEYS = ("endpoints", "endcolors", "range", "endrem", "tri_id")


def _np(out):
  return {k: v.cpu().numpy() for k, v in out.items()}


def _assert_bit_equal(a, b, keys=KEYS):
  for k in keys:
    x, y = np.asarray(a[k]), np.asarray(b[k])
    assert x.shape == y.shape, k
    same = x.view(np.int32) == y.view(np.int32)
    assert same.all(), "%s: %d of %d differ (first at %d: %r vs %r)" % (
        k, (~same).sum(), same.size, np.flatnonzero(~same.reshape(-1))[0],
        x.reshape(-1)[np.flatnonzero(~same.reshape(-1))[0]], y.reshape(-1)[np.flatnonzero(~same.reshape(-1))[0]])


@pytest.mark.parametrize("n_side,H,W", [(40, 8, 64), (120, 64, 512), (300, 64, 2048)])
def test_bvh_trace_bit_exact_vs_oracle(engine, oracle, n_side, H, W):
  sc = synth.make_scene(1000 + n_side, n_side=n_side)
  rays = oracle.create_rays(3.0, -25.0, H, W)
  origin = np.zeros(3, np.float32)
  ref = oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
  st = bvh.status()
  assert st["n_tris"] == sc["faces"].shape[0] and st["n_bad_faces"] == 0
  got = _np(engine.trace(bvh, rays, origin, H))
  assert (got["tri_id"] >= 0).mean() > 0.9
  _assert_bit_equal(got, ref)


def test_bruteforce_kernel_matches_oracle_bruteforce(engine, oracle):
  sc = synth.make_scene(7, n_side=40)
  rays = oracle.create_rays(10.0, -30.0, 16, 64)
  origin = np.array([0.5, -0.25, 0.3], np.float32)
  ref = oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], 16, oracle.BRUTE_FORCE | oracle.NORMALIZE_SSE)
  got = _np(engine.trace_bruteforce(sc["verts"], sc["faces"], sc["colors"], sc["rem"], rays, origin, 16))
  _assert_bit_equal(got, ref)
  bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
  got2 = _np(engine.trace(bvh, rays, origin, 16))
  _assert_bit_equal(got2, ref)


def test_host_ctrace_drop_in(engine, oracle):
  sc = synth.make_scene(11, n_side=60)
  H, W = 16, 128
  rays = oracle.create_rays(3.0, -25.0, H, W)
  origin = np.zeros(3, np.float32)
  ref = oracle.trace(rays, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  out = dict(endpoints=np.full(3 * H * W, 7.0, np.float32), endcolors=np.full(3 * H * W, 7, np.int32),
             range=np.full(H * W, 7.0, np.float32), endrem=np.full(H * W, 7.0, np.float32))
  # add rays that miss (pointing up) to check that misses leave the buffers untouched
  rays2 = rays.copy()
  rays2[:W] = np.array([0, 0, 1], np.float32)
  ref2 = oracle.trace(rays2, origin, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  got = engine.ctrace_host(rays2, origin, sc["verts"].reshape(-1), sc["faces"].reshape(-1), sc["colors"].reshape(-1),
                           sc["rem"], H, outputs=out, want_ids=True)
  miss = ref2["tri_id"] < 0
  assert miss[:W].all()
  assert (got["range"][miss] == 7.0).all() and (got["endcolors"].reshape(-1, 3)[miss] == 7).all()
  hit = ~miss
  assert np.array_equal(got["tri_id"], ref2["tri_id"])
  assert np.array_equal(got["range"][hit].view(np.int32), ref2["range"][hit].view(np.int32))
  assert np.array_equal(got["endpoints"].reshape(-1, 3)[hit].view(np.int32), ref2["endpoints"].reshape(-1, 3)[hit].view(np.int32))
  assert np.array_equal(got["endcolors"].reshape(-1, 3)[hit], ref2["endcolors"].reshape(-1, 3)[hit])


@pytest.mark.parametrize("n_side,H,W,origin", [(40, 8, 64, (0, 0, 0)), (120, 64, 512, (0.5, -0.25, 0.3)), (300, 64, 2048, (0, 0, 0)),
                                                  (60, 3, 50, (0, 0, 0)), (2, 4, 8, (0, 0, 0))])
def test_persistent_tma_trace_is_bit_identical(engine, oracle, n_side, H, W, origin):
  """North-star (ii) as specified -- persistent warps pulling rays from a counter, warp-wide ray compaction, the top 11
  levels of the tree staged in shared memory by cp.async.bulk + mbarrier (k_trace_persistent, VL_TRACE_PERSISTENT) --
  against the oracle and against the default per-ray kernel: beam grids, a ray set that is not a grid (3 x 50), a
  two-triangle mesh whose root is a leaf, ragged tiles, rays that miss."""
  sc = synth.make_scene(2000 + n_side, n_side=max(n_side, 2), n_boxes=4 if n_side > 10 else 0)
  rays = oracle.create_rays(10.0, -30.0, H, W)
  rays[::7] = np.array([0.0, 0.0, 1.0], np.float32)      # straight up: misses, and zero direction components
  o = np.asarray(origin, np.float32)
  ref = oracle.trace(rays, o, sc["verts"], sc["faces"], sc["colors"], sc["rem"], H, oracle.MIN_ID_TIES | oracle.NORMALIZE_SSE)
  bvh = engine.Bvh(sc["verts"], sc["faces"], sc["colors"], sc["rem"])
  a = _np(engine.trace(bvh, rays, o, H, persistent=True))
  b = _np(engine.trace(bvh, rays, o, H))
  _assert_bit_equal(a, b)
  _assert_bit_equal(a, ref)
  z = _np(engine.trace(bvh, rays, o, H, persistent=True, zero_misses=True, out={k: v for k, v in engine.trace(bvh, rays, o, H).items()}))
  miss = ref["tri_id"] < 0
  assert (z["range"][miss] == 0).all() and np.array_equal(z["tri_id"], ref["tri_id"])
