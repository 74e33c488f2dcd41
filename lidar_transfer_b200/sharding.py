"""Scan sharding across GPUs: one process per GPU, scans are independent units, NO collective on the data
path (BASELINE.json north_star).  torch.distributed is used only to agree on timing
(barrier / max-over-ranks) and to gather per-rank statistics for the report."""
import os

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


def _cpulist(text):
  cpus = set()
  for part in text.strip().split(","):
    if not part:
      continue
    a, _, b = part.partition("-")
    cpus.update(range(int(a), int(b or a) + 1))
  return cpus


def bind_to_gpu_numa_node(device_index):
  """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node and the
  node's cpulist), so that the pinned host buffers it allocates afterwards (first touch) and the thread that feeds
  the GPU sit next to the PCIe root port.  Host meshes cross PCIe once per scan -- with 8 ranks on one box the
  end-to-end path is bound by host memory / PCIe topology, not by the device.  Returns a short description of what
  was done (also when nothing could be done: containers often hide the topology)."""
  try:
    import torch
    bus = torch.cuda.get_device_properties(device_index).pci_bus_id
    dom = torch.cuda.get_device_properties(device_index).pci_domain_id
    dev = torch.cuda.get_device_properties(device_index).pci_device_id
    addr = "%04x:%02x:%02x.0" % (dom, bus, dev)
    node = int(open("/sys/bus/pci/devices/%s/numa_node" % addr).read().strip())
    if node < 0:
      return "numa node of %s unknown" % addr
    want = _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
    have = os.sched_getaffinity(0)
    use = want & have
    if not use:
      return "no allowed cpu on numa node %d of %s" % (node, addr)
    os.sched_setaffinity(0, use)
    return "bound to %d cpus of numa node %d (%s)" % (len(use), node, addr)
  except (OSError, ValueError, AttributeError, RuntimeError) as e:
    return "not bound (%s)" % e
