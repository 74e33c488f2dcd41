#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.

Extracts the reference's CUDA `integrate` kernel source -- the string handed to
pycuda's SourceModule at auxiliary/fusion_lidar.py:66-229 -- from the reference
checkout where it lies and writes it, unmodified, to oracle/_ref/integrate_kernel.inc
(git-ignored) so that oracle/ref_tsdf_harness.cpp can compile it for the CPU.
No reference source is copied into the repository history.
"""
import re
import sys

src_path, out_path = sys.argv[1], sys.argv[2]
text = open(src_path).read()
m = re.search(r'self\._cuda_src_mod = SourceModule\("""(.*?)"""\)', text, re.S)
if m is None:
  sys.exit("integrate kernel string not found in %s" % src_path)
kernel = m.group(1)
assert "__global__ void integrate" in kernel
open(out_path, "w").write(kernel + "\n")
print("extracted %d bytes of kernel source -> %s" % (len(kernel), out_path))
