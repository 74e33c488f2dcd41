"""Wall time of the drop-in call the reference's driver makes per scan (lidar_deform.py:393-452): scan files on disk ->
MultiSemLaserScan.open_multiple_scans -> deform('mergemesh') -> result arrays, at config-1 size (124 668 points,
64x2048 source and target, voxel 0.05 m, bounds of config/lidar_transfer.yaml clipped to the points like the
reference does).  usage: deform_bench.py [voxel] [n_calls]"""
import cProfile, io, json, os, pstats, sys, tempfile, time, contextlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_transfer_b200 import synth
from lidar_transfer_b200.auxiliary.laserscan import MultiSemLaserScan

vox = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
n_calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tmp = tempfile.mkdtemp()
names, labels = [], []
for k in range(2):
  pts, lab = synth.make_scan_points(10 + k, 124668)
  p, l = os.path.join(tmp, "%06d.bin" % k), os.path.join(tmp, "%06d.label" % k)
  pts.astype(np.float32).tofile(p); lab.astype(np.uint32).tofile(l)
  names.append(p); labels.append(l)
lut = {0: [0, 0, 0]}
for v in synth.STATIC_LABELS:
  lut[int(v)] = [int(v) % 251 + 1, (7 * int(v)) % 256, 9]
src = dict(name="HDL-64E", beams=64, fov_up=3.0, fov_down=-25.0, fov_hor=360.0, angle_res_hor=360.0 / 2048)
poses = [np.eye(4), np.eye(4)]
res = {"voxel": vox}
times = []
prof = cProfile.Profile()
for i in range(n_calls):
  ms = MultiSemLaserScan(src, src, 1, max(lut) + 1, [0, 1], [252, 259], lut, transformation=None, preserve_float=True,
                         voxel_size=vox, vol_bnds=np.array([[-50, 50], [-50, 50], [-3, 2]]))
  with contextlib.redirect_stdout(io.StringIO()):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms.open_multiple_scans(names, labels, poses, 0)
    t1 = time.perf_counter()
    if i == n_calls - 1: prof.enable()
    out = ms.deform('mergemesh', poses, 0)
    if i == n_calls - 1: prof.disable()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
  times.append((t1 - t0, t2 - t1))
res["open_ms"] = [round(1e3 * a, 1) for a, b in times]
res["deform_ms"] = [round(1e3 * b, 1) for a, b in times]
res["n_faces"] = int(out[2].shape[0])
res["hit_fraction"] = float((ms.proj_range > 0).mean())
s = io.StringIO()
pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(18)
print(s.getvalue()[-3500:], file=sys.stderr)
print(json.dumps(res))
