import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
  """The CPU checker (test infrastructure): builds liboracle.so on demand, and oracle/_ref when the
  reference checkout is present (otherwise the prebuilt oracle/_ref binaries are used if shipped)."""
  from oracle import oracle as O
  O.build()
  return O


@pytest.fixture(scope="session")
def vl():
  """The product library, built in-tree on demand."""
  from lidar_transfer_b200 import build
  build.build()
  from lidar_transfer_b200 import _lib
  return _lib.lib()


@pytest.fixture(scope="session")
def engine(vl):
  import torch
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  from lidar_transfer_b200 import engine as E
  return E


@pytest.fixture(autouse=True)
def _poison_recycled_device_memory(request):
  """VL_POISON=1 (stress aid): before every GPU test a large block of device memory is filled with 0xCD and handed back
  to torch's caching allocator, so that every `torch.empty` the test makes comes out of memory full of garbage -- a
  kernel or a status read that relies on memory it never initialised then fails here instead of once in a while."""
  if os.environ.get("VL_POISON") and request.node.get_closest_marker("gpu") is not None:
    import torch
    if torch.cuda.is_available():
      torch.cuda.empty_cache()
      junk = [torch.full((1 << 30,), 0xCD, dtype=torch.uint8, device="cuda") for _ in range(6)]
      del junk
      torch.cuda.synchronize()
  yield
