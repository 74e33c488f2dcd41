"""Makes `import auxiliary...` resolve to this package's mirror of the reference's `auxiliary` package, so that code
written against the reference -- above all its own driver, lidar_deform.py (`from auxiliary.laserscan import *`,
`from auxiliary.laserscanvis import LaserScanVis`, lidar_deform.py:9-10) -- runs UNCHANGED on libvlidar.

    import lidar_transfer_b200.dropin as dropin; dropin.install()          # in-process
    python -m lidar_transfer_b200.dropin /path/to/lidar_deform.py -d minimal -b -w -o out    # run a driver script

The modules are registered in sys.modules under the reference's names (the package's relative imports keep
working, which a bare sys.path entry would break).  install() refuses to shadow a real `auxiliary` package that was
imported before."""
import importlib
import runpy
import sys

_NAMES = ("", ".laserscan", ".fusion_lidar", ".raytracing", ".np_ioueval", ".laserscanvis", ".raytracer",
          ".raytracer.RayTracerCython")


def install():
  mine = importlib.import_module("lidar_transfer_b200.auxiliary")
  have = sys.modules.get("auxiliary")
  if have is not None and have is not mine:
    raise ImportError("a different `auxiliary` package is already imported from %s" % getattr(have, "__file__", "?"))
  for suffix in _NAMES:
    sys.modules["auxiliary" + suffix] = importlib.import_module("lidar_transfer_b200.auxiliary" + suffix)
  return mine


def run_driver(script, argv):
  """Runs a driver script (source or compiled .pyc) as __main__ with `auxiliary` mapped to this package.  The
  reference's driver leaves through quit(): SystemExit is the normal end of a batch (lidar_deform.py:452-459)."""
  install()
  old = sys.argv
  sys.argv = [script] + list(argv)
  try:
    runpy.run_path(script, run_name="__main__")
  except SystemExit as e:
    if e.code not in (None, 0):
      raise
  finally:
    sys.argv = old


if __name__ == "__main__":
  if len(sys.argv) < 2:
    sys.exit("usage: python -m lidar_transfer_b200.dropin <driver.py> [driver arguments]")
  run_driver(sys.argv[1], sys.argv[2:])
