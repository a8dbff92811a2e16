import time, numpy as np, sys
sys.path.insert(0, '.')
import edt_b200
rng = np.random.default_rng(0)
lab = np.asfortranarray(rng.integers(0, 256, (512,)*3, dtype=np.uint32))
edt_b200.edtsq(lab)
t0 = time.perf_counter()
for _ in range(3): out = edt_b200.edtsq(lab)
print("numpy (pageable) edtsq 512^3: %.1f ms per call" % ((time.perf_counter() - t0) / 3 * 1e3))
