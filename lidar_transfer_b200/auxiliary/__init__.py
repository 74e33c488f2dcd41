"""Host-side mirror of the reference's `auxiliary` package for the hot path.

Same module names, function names, argument order and error behaviour as
/root/reference/auxiliary (PRBonn/lidar_transfer), so a pipeline written against the reference
(`lidar_deform.py`: `from auxiliary.laserscan import *`) runs on libvlidar by putting
`lidar_transfer_b200` first on sys.path -- see INTEGRATION.md.  All compute goes through the C ABI
in include/vlidar.h; there is no CPU fallback.
"""
