import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
  """The oracle is test infrastructure: build it on demand (g++ only, a few seconds)."""
  import subprocess
  so = os.path.join(ROOT, "oracle", "liboracle.so")
  src = os.path.join(ROOT, "oracle", "oracle.cc")
  if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
  yield
