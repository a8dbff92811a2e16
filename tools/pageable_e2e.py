"""numpy (pageable) end-to-end time of edtsq on 512^3 uint32: what a drop-in Python user sees.
Knob (environment): EDTB200_COPY_THREADS, the number of host threads filling / draining the
pinned staging buffers (default 16 on hosts with >= 32 hardware threads)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edt_b200  # noqa: E402

rng = np.random.default_rng(0)
lab = np.asfortranarray(rng.integers(0, 256, (512,) * 3, dtype=np.uint32))
edt_b200.edtsq(lab)
times = []
for _ in range(5):
  t0 = time.perf_counter()
  out = edt_b200.edtsq(lab)
  times.append((time.perf_counter() - t0) * 1e3)
  del out
print("numpy (pageable) edtsq 512^3 uint32: min %.1f ms, median %.1f ms per call  [copy threads %s]" % (
  min(times), sorted(times)[len(times) // 2], os.environ.get("EDTB200_COPY_THREADS", "default")))
