import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
  try:
    import edt_b200
    return edt_b200.device_count() > 0
  except Exception:
    return False


def pytest_collection_modifyitems(config, items):
  if _has_gpu():
    return
  skip = pytest.mark.skip(reason="no CUDA device visible")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
  from oracle import oracle as mod
  mod._lib()
  return mod


@pytest.fixture(scope="session")
def reference(oracle):
  """The compiled, unmodified reference (oracle/_ref) or None when it has not been built."""
  return oracle.load_reference()


@pytest.fixture(scope="session")
def edt():
  import edt_b200
  edt_b200._lib()
  return edt_b200
