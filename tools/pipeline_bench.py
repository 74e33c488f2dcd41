"""The sharded real-pipeline leg of bench.py alone (points -> image -> 284 M-voxel TSDF -> mesh -> cast -> results), one GPU:
    python tools/pipeline_bench.py [n_scans=300] [lanes=1,2,3]      scans/s per number of scans in flight"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
lanes = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,2,3").split(",")]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
res = {}
for L in lanes:
  os.environ["VL_PIPE_LANES"] = str(L)
  r = bench.sharded_pipeline_leg(0, 1, dev, n, torch.cuda.synchronize, lambda ms: ms)
  res["lanes_%d" % L] = dict(scans_per_s=round(n / (r["ms"] * 1e-3), 1), ms_per_scan=round(r["ms"] / n, 4), hit_fraction=r["hit_fraction"])
  torch.cuda.empty_cache()
print(json.dumps(res))
