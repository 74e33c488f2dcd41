"""Placeholder for the reference's `auxiliary.laserscanvis` (the vispy GUI, auxiliary/laserscanvis.py): out of scope
 -- the module exists so that the reference's driver, which imports it unconditionally
(lidar_deform.py:10), loads; it only instantiates LaserScanVis without -b / --batch (lidar_deform.py:365-377)."""


class LaserScanVis:
  def __init__(self, *args, **kwargs):
    raise NotImplementedError("the vispy visualiser is not part of lidar_transfer_b200: run the driver with -b / --batch")
