#!/usr/bin/env python
"""A few small transforms that touch every kernel variant, for runs under compute-sanitizer:
  compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
  compute-sanitizer --tool racecheck python tools/sanitize_cases.py
Results are still compared with the oracle, so a sanitizer-induced slowdown cannot hide a wrong answer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import edt_b200 as edt  # noqa: E402
from oracle import oracle  # noqa: E402

rng = np.random.default_rng(3)
done = 0


def check(name, got, want):
  global done
  assert np.array_equal(got, want, equal_nan=True), name
  done += 1


# vector X pass + TMA tile kernel (sx % 4 == 0), short / long runs, all epilogues
for kind in ("iid", "blocks", "balls"):
  lab = np.asfortranarray(cases.random_volume(rng, (64, 80, 72), kind, np.uint32))
  check("edtsq " + kind, edt.edtsq(lab, anisotropy=(1, 2, 3)), oracle.edtsq(lab, anisotropy=(1, 2, 3)))
  check("sdf " + kind, edt.sdf(lab, anisotropy=(1, 2, 3), black_border=True),
        oracle.sdf(lab, anisotropy=(1, 2, 3), black_border=True))
# generic X pass + plain-load tile kernel (odd sizes), every label width
for dtype in (np.uint8, np.uint16, np.uint64):
  lab = np.asfortranarray(cases.random_volume(rng, (37, 45, 51), "blocks", dtype))
  check("odd " + np.dtype(dtype).name, edt.edt(lab, anisotropy=(0.7, 1.3, 2.9)), oracle.edt(lab, anisotropy=(0.7, 1.3, 2.9)))
# long runs across many chunks (stitching stage), 2-D and 1-D drivers
lab = np.ones((8, 300, 12), dtype=np.uint8, order="F")
lab[3, 150, 5] = 0
check("long runs", edt.edtsq(lab), oracle.edtsq(lab))
img = cases.random_volume(rng, (200, 130), "balls", np.uint16)
check("2-D", edt.edt(img, black_border=True), oracle.edt(img, black_border=True))
row = cases.random_volume(rng, (700,), "blocks", np.uint32)
check("1-D", edt.edtsq(row, anisotropy=2.5), oracle.edtsq(row, anisotropy=2.5))
# lines beyond the tile kernel (n > 4096): thread-per-line kernel
tall = np.asfortranarray(cases.random_volume(rng, (8, 4200, 2), "blocks", np.uint8))
check("n > 4096", edt.edtsq(tall), oracle.edtsq(tall))
# voxel graph
lab, graph, kw = cases.random_graph_case(7)
check("voxel graph", edt.edtsq(lab, voxel_graph=graph, **kw), oracle.edtsq(lab, voxel_graph=graph, **kw))
# the 2-CTA variant of the tile kernel (label noise seen by the previous transform) and the double
# hull tests (non-integer weights) on long varying runs; a Voronoi-like volume on a 512-row axis
lab = np.asfortranarray(cases.random_volume(rng, (64, 80, 72), "iid", np.uint32))
for _ in range(2):
  got = edt.edtsq(lab)
check("noise twice", got, oracle.edtsq(lab))
from scipy.spatial import cKDTree  # noqa: E402
pts = rng.uniform(0, 1, (12, 3)) * np.array([40, 512, 24])
grid = np.stack(np.meshgrid(np.arange(40.0), np.arange(512.0), np.arange(24.0), indexing="ij"), -1).reshape(-1, 3)
vor = np.asfortranarray((cKDTree(pts).query(grid)[1] + 1).astype(np.uint16).reshape(40, 512, 24))
check("voronoi int", edt.edtsq(vor, anisotropy=(1, 1, 2)), oracle.edtsq(vor, anisotropy=(1, 1, 2)))
check("voronoi double", edt.edtsq(vor, anisotropy=(0.7, 1.3, 2.9)), oracle.edtsq(vor, anisotropy=(0.7, 1.3, 2.9)))
# wide variant (lines of 513..1024 rows: decoupled warps, chunk mask) and the constant runs that are
# not stored: solid volume with a black border through every epilogue, blocks, noise, one hole
wide = np.ones((16, 1024, 8), dtype=np.uint8, order="F")
check("wide ones bb", edt.edtsq(wide, anisotropy=(6, 6, 30), black_border=True),
      oracle.edtsq(wide, anisotropy=(6, 6, 30), black_border=True))
check("wide ones sdf", edt.sdf(wide, black_border=True), oracle.sdf(wide, black_border=True))
wide[7, 600, 3] = 0
check("wide hole", edt.edt(wide), oracle.edt(wide))
for kind in ("iid", "blocks"):
  lab = np.asfortranarray(cases.random_volume(rng, (16, 1000, 12), kind, np.uint32))
  check("wide " + kind, edt.edtsq(lab, black_border=(kind == "blocks")), oracle.edtsq(lab, black_border=(kind == "blocks")))
solid = np.ones((48, 96, 40), dtype=np.uint16, order="F")
solid[:, :, 20:] = 2
check("solid edt", edt.edt(solid, anisotropy=(2, 1, 3), black_border=True),
      oracle.edt(solid, anisotropy=(2, 1, 3), black_border=True))
# per-label statistics and box-restricted extraction
import torch  # noqa: E402
lt = torch.from_numpy(np.ascontiguousarray(vor.astype(np.int32))).cuda()
dt = edt.edt_cuda(lt, (1.0, 1.0, 1.0), False, sqrt=True)
st = edt.label_stats_cuda(lt, dt)
assert int(st["count"].sum().item()) == lt.numel()
for key, img in edt.each_cuda(lt, dt, in_place=True):
  assert torch.equal(img, torch.where(lt == key, dt, torch.zeros((), device=dt.device)))
done += 1
print("sanitize_cases: %d results equal to the oracle" % done)
