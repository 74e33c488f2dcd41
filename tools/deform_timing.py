"""Where the wall time of the reference-shaped per-scan call goes: open_multiple_scans + deform('mergemesh') on the real
fixture at config-1 size, with the device time of the chain beside it.  usage: deform_timing.py [n_rounds]"""
import cProfile, io, json, os, pstats, sys, tempfile, time, zipfile, contextlib
import numpy as np, torch, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_transfer_b200.auxiliary import laserscan as ls

d = tempfile.mkdtemp()
zipfile.ZipFile(os.path.join(ROOT, "tests", "golden", "minimal_fixture.zip")).extractall(d)
cfg = yaml.safe_load(open(d + "/config/lidar_transfer.yaml")); src = yaml.safe_load(open(d + "/minimal/config.yaml"))
seq = d + "/minimal/sequences/00"
sn = [seq + "/velodyne/%06d.bin" % k for k in range(3)]; ln = [seq + "/labels/%06d.label" % k for k in range(3)]
poses = [np.eye(4)] * 3
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rows = []
prof = cProfile.Profile()
for r in range(rounds):
  for idx in range(3):
    with contextlib.redirect_stdout(io.StringIO()):
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      scans = ls.MultiSemLaserScan(src, src, 1, len(cfg["color_map"]), cfg["ignore"], cfg["moving"], cfg["color_map"],
                                   transformation=cfg["transformation"], preserve_float=True, voxel_size=0.05,
                                   vol_bnds=np.array(cfg["voxel_bounds"]).reshape(3, 2))
      scans.open_multiple_scans(sn, ln, poses, idx)
      t1 = time.perf_counter()
      if r == rounds - 1: prof.enable()
      scans.deform("mergemesh", poses, idx)
      t2 = time.perf_counter()
      x = scans.proj_range[0, 0] + scans.label_image[0, 0] + scans.back_points[0, 0] + scans.proj_remissions[0, 0]
      if r == rounds - 1: prof.disable()
      t3 = time.perf_counter()
    rows.append((r, idx, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
for row in rows: print("round %d scan %d: open %.2f ms, deform %.2f ms, results to host %.2f ms" % row)
s = io.StringIO(); pstats.Stats(prof, stream=s).sort_stats("tottime").print_stats(16); print(s.getvalue()[-2800:])
