"""The three index tricks the scene-streaming cast added in round 2 (lidar_transfer_b200/csrc/vl_cast.cu), restated in numpy
float32 and checked on the CPU against their definitions -- each of them may only REMOVE work that cannot produce a hit:

  beam_in        "is there a beam whose sine lies in [slo, shi]?" answered by one load from the next-beam-sine table
                 (nxt[b] = smallest beam sine among the fine bins >= b): it may say yes too often, never no wrongly;
  row trim       a rectangle drops the cell rows at both ends whose [min, max] beam sine misses [slo, shi]: the rows that
                 remain contain every beam the per-beam filter of k_cast_units would let through;
  unit of a candidate   k_cast_units finds the unit of a pooled candidate from one word of start marks
                 (popc(marks & bits 0 .. lane)) instead of a bisection of the unit offsets: the same unit, always.

The GPU tests check the consequence (bit-identical hits vs the oracle); these cover the arguments themselves over ray sets and
intervals nobody enumerated (grid sensors, random sets, degenerate ranges)."""
import numpy as np
from hypothesis import given, settings, strategies as st

F = np.float32
K_FINE = 4096


def _params(sines, ch):
  lo, hi = F(sines.min()), F(sines.max())
  span = F(hi - lo)
  ok = span > F(1e-12)
  return lo, hi, (F(ch) / span if ok else F(0)), (F(K_FINE) / span if ok else F(0))


def _fine_of(s, lo, nf_inv):
  return np.clip(np.floor((np.asarray(s, F) - lo) * nf_inv).astype(np.int64), 0, K_FINE - 1)


def _row_of(s, lo, ch_inv, ch):
  return np.clip(np.floor((np.asarray(s, F) - lo) * ch_inv).astype(np.int64), 0, ch - 1)


def _beam_sets(seed):
  rng = np.random.default_rng(seed)
  kind = seed % 4
  if kind == 0:      # a grid sensor: H rows, every row's sine repeated with an ulp of jitter (the host normaliser's)
    H = int(rng.integers(2, 129))
    pitch = np.deg2rad(np.linspace(rng.uniform(0, 25), rng.uniform(-35, -5), H))
    s = np.repeat(np.sin(pitch).astype(F), 16)
    s = np.nextafter(s, F(2)) if seed % 8 < 4 else s
    return s.astype(F), H
  if kind == 1:      # random directions
    n = int(rng.integers(1, 3000))
    return rng.uniform(-1, 1, n).astype(F), int(rng.integers(1, 200))
  if kind == 2:      # two tight clusters far apart
    n = int(rng.integers(2, 500))
    return np.concatenate([rng.normal(-0.4, 1e-6, n), rng.normal(0.05, 1e-6, n)]).astype(F), 64
  return np.full(int(rng.integers(1, 50)), F(rng.uniform(-0.9, 0.9)), F), 8      # one sine only: empty range


@settings(max_examples=120, deadline=None)
@given(st.integers(0, 10 ** 6))
def test_beam_in_never_says_no_wrongly_and_row_trim_keeps_every_beam_in_the_interval(seed):
  sines, ch = _beam_sets(seed)
  lo, hi, ch_inv, nf_inv = _params(sines, ch)
  # the table k_beam_count / k_beam_scan_top build: per fine bin the smallest sine, then the suffix minimum
  per_bin = np.full(K_FINE, np.inf, F)
  np.minimum.at(per_bin, _fine_of(sines, lo, nf_inv), sines)
  nxt = np.minimum.accumulate(per_bin[::-1])[::-1]
  rows = _row_of(sines, lo, ch_inv, ch)
  rmin, rmax = np.full(ch, np.nan, F), np.full(ch, np.nan, F)
  for r in np.unique(rows):
    rmin[r], rmax[r] = sines[rows == r].min(), sines[rows == r].max()
  rng = np.random.default_rng(seed + 1)
  centre = np.concatenate([rng.choice(sines, 200), rng.uniform(-1.2, 1.2, 200)]).astype(F)
  half = (10.0 ** rng.uniform(-7, -0.5, centre.size)).astype(F)
  los, his = (centre - half).astype(F), (centre + half).astype(F)
  # ... and intervals that END exactly on a beam's sine (the comparisons are inclusive, like the per-beam filter's)
  e0, e1 = rng.choice(sines, 100), rng.choice(sines, 100)
  los = np.concatenate([los, np.minimum(e0, e1), e0]).astype(F)
  his = np.concatenate([his, np.maximum(e0, e1), e0]).astype(F)
  for slo, shi in zip(los, his):
    inside = (sines >= slo) & (sines <= shi)               # what the per-beam filter lets through
    if shi < lo or slo > hi:                               # tri_setup returns before either trick
      assert not inside.any()
      continue
    beam_in = nxt[_fine_of(slo, lo, nf_inv)] <= shi
    assert beam_in or not inside.any()
    if not beam_in:
      continue
    ra, rb = int(_row_of(slo, lo, ch_inv, ch)), int(_row_of(shi, lo, ch_inv, ch))
    with np.errstate(invalid="ignore"):
      while ra <= rb and not (rmax[ra] >= slo and rmin[ra] <= shi):
        ra += 1
      while rb > ra and not (rmax[rb] >= slo and rmin[rb] <= shi):
        rb -= 1
    kept = (rows >= ra) & (rows <= rb) if ra <= rb else np.zeros(sines.size, bool)
    assert not (inside & ~kept).any(), (slo, shi, ra, rb)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 10 ** 6))
def test_candidate_to_unit_by_start_marks_equals_the_bisection(seed):
  rng = np.random.default_rng(seed)
  n_r = rng.integers(0, 20, 32) * (rng.random(32) < rng.uniform(0.2, 1.0))       # beams per unit of one warp's group, many empty
  n_r = n_r.astype(np.int64)
  off = np.cumsum(n_r) - n_r
  total = int(n_r.sum())
  nz = n_r > 0
  rank_of_lane = np.cumsum(nz) - nz                                              # slot of a unit among the parked ones
  parked_lane = np.flatnonzero(nz)                                               # slot -> lane
  jb = 0
  for base in range(0, total, 32):
    rel = off - base
    marks = 0
    for lane in range(32):
      if nz[lane] and 0 <= rel[lane] < 32:
        marks |= 1 << int(rel[lane])
    for lane in range(32):
      it = base + lane
      if it >= total:
        break
      slot = jb + bin(marks & (0xFFFFFFFF >> (31 - lane))).count("1") - 1
      want = int(np.searchsorted(off + n_r, it, side="right"))                   # the unit whose range holds candidate `it`
      assert parked_lane[slot] == want and rank_of_lane[want] == slot
      assert 0 <= it - off[want] < n_r[want]
    jb += bin(marks).count("1")
