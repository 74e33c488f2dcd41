"""Builds tests/golden/minimal_fixture.zip from the reference's own fixture (/root/reference/minimal.zip, SURVEY.md
section 2 #16: three KITTI sequence-00 scans with labels, poses, calibration, the HDL-64E source config and the
HDL-32E target config) plus the reference's approach configuration (config/lidar_transfer.yaml, consumed as-is) and
an OS1-128 target for BASELINE.json configs[3].  DATA only -- no reference code.  Run in the authoring container:

    python tests/golden/make_minimal_fixture.py

The GPU box has no /root/reference; tests unpack this archive into a temporary directory."""
import io
import os
import zipfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "minimal_fixture.zip")

OS1_128 = """name: "Ouster OS1-128 (BASELINE.json configs[3] target)"
beams: 128
fov_up: 22.5
fov_down: -22.5
fov_hor: 360
angle_res_hor: 0.17578125 # 2048
res_hor: 2048
height: 1.73
beam_angles:
"""


def main():
  src = zipfile.ZipFile(os.path.join(REF, "minimal.zip"))
  with zipfile.ZipFile(OUT, "w", zipfile.ZIP_DEFLATED, compresslevel=9) as z:
    for info in sorted(src.infolist(), key=lambda i: i.filename):
      if info.is_dir():
        continue
      zi = zipfile.ZipInfo(info.filename, date_time=(2020, 1, 1, 0, 0, 0))   # fixed time stamps: reproducible bytes
      zi.compress_type = zipfile.ZIP_DEFLATED
      z.writestr(zi, src.read(info.filename))
    for name, data in (("config/lidar_transfer.yaml", open(os.path.join(REF, "config", "lidar_transfer.yaml"), "rb").read()),
                       ("config/os1_128.yaml", OS1_128.encode())):
      zi = zipfile.ZipInfo(name, date_time=(2020, 1, 1, 0, 0, 0))
      zi.compress_type = zipfile.ZIP_DEFLATED
      z.writestr(zi, data)
  print(OUT, os.path.getsize(OUT), "bytes", len(zipfile.ZipFile(OUT).namelist()), "files")


if __name__ == "__main__":
  main()
