import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_cuda_library():
  """Build libmono_emb.so when a checkout arrives without it (the .so is git-ignored; nvcc cross-compiles in ~2 min).
  This is the same recipe as __graft_entry__.build(); a failed build fails the session loudly — there is no fallback."""
  import glob
  import subprocess
  so = os.path.join(ROOT, "monolith_b200", "lib", "libmono_emb.so")
  srcs = glob.glob(os.path.join(ROOT, "monolith_b200", "csrc", "*.cu*")) + \
      glob.glob(os.path.join(ROOT, "monolith_b200", "csrc", "*.h")) + [os.path.join(ROOT, "include", "mono_emb.h")]
  if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
    subprocess.check_call(["bash", os.path.join(ROOT, "monolith_b200", "csrc", "build.sh")])
  yield


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
  """The oracle is test infrastructure: build it on demand (g++ only, a few seconds)."""
  import subprocess
  so = os.path.join(ROOT, "oracle", "liboracle.so")
  src = os.path.join(ROOT, "oracle", "oracle.cc")
  if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
  yield
