"""Scan sharding across GPUs: one process per GPU, scans are independent units, NO collective on the data
path (BASELINE.json north_star).  torch.distributed is used only to agree on timing
(barrier / max-over-ranks) and to gather per-rank statistics for the report."""
import torch.distributed as dist


def scans_for_rank(n_scans, rank, world):
  """Round-robin assignment scan k -> rank k mod world (the reference's driver walks scans serially,
  lidar_deform.py:393-458; any partition is valid because scans share nothing)."""
  return list(range(rank, n_scans, world))


def gather_stats(stats):
  """All ranks' stats dicts, in rank order (host-side object gather, off the data path)."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
    return [stats]
  out = [None] * dist.get_world_size()
  dist.all_gather_object(out, stats)
  return out


def aggregate(per_rank, rays_per_scan):
  """Whole-job throughput: all scans of all ranks / the slowest rank's time."""
  scans = sum(s["n"] for s in per_rank)
  ms = max(s["ms"] for s in per_rank)
  return {"scans": scans, "ms": ms, "scans_per_s": scans / (ms * 1e-3), "mrays_per_s": scans * rays_per_scan / (ms * 1e-3) / 1e6}
