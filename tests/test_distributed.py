"""CPU tests (gloo, world_size 2 and 3) of the slab-split host logic in
euclidean-distance-transform-3d_b200/distributed.py.

There is no GPU here, so the per-axis device entry points are replaced by the oracle's per-axis
passes (test infrastructure) -- what is under test is the decomposition itself: slab geometry,
the Z-slab <-> Y-slab exchange, border flags, uneven and empty slabs.  The same code runs over
NCCL in the `-m gpu` test below and in bench.py --gpus N.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OraclePasses:
  """Stand-in for CudaPasses: same interface, numpy arithmetic from oracle/ (tests only)."""

  def __init__(self):
    from oracle import oracle
    self.oracle = oracle
    self.signed = False

  def empty_f32(self, shape):
    return torch.zeros(shape, dtype=torch.float32)

  def pass_first(self, labels, f, wx, black_border, signed):
    lab = labels.numpy()
    self.signed = bool(signed)
    if signed:                      # background as an ordinary label == shift every label by one
      lab = lab.astype(np.int64) + 1
    f.copy_(torch.from_numpy(self.oracle.pass_first(lab, wx, black_border)))

  def pass_later(self, labels, f, axis, w, border_lo, border_hi, sqrt=False, negate=False):
    lab = labels.numpy().astype(np.int64)
    zero = lab == 0
    arr = f.numpy()
    self.oracle.pass_later(lab + 1, arr, axis, w, border_lo, border_hi)   # +1: every run is foreground...
    if not self.signed:
      # ...except that plain EDT keeps background at 0 (its rows hold 0 already and stay 0)
      arr[zero] = 0.0
    if sqrt:
      np.sqrt(arr, out=arr)
    if negate:
      arr[zero] *= -1.0


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _volume(case):
  shape, kind = case[0], case[1]
  rng = np.random.default_rng(1234)
  if kind == "iid":
    vol = rng.integers(0, 4, shape)
  elif kind == "ones":
    vol = np.ones(shape, dtype=np.int64)
    vol[tuple(s // 2 for s in shape)] = 0
  else:
    small = rng.integers(0, 3, tuple((s + 4) // 5 for s in shape))
    vol = np.repeat(np.repeat(np.repeat(small, 5, 0), 5, 1), 5, 2)[:shape[0], :shape[1], :shape[2]]
  return np.ascontiguousarray(vol.astype(np.int32))


def _worker(rank, world, port, queue):
  sys.path.insert(0, ROOT)
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    import edt_b200.distributed as ed
    for idx, case in enumerate(CASES):
      shape, kind, an, bb, sqrt, signed, depths = case
      if depths is not None and len(depths) != world:
        continue
      vol = _volume(case)
      if depths is None:
        parts = ed.split_extent(shape[0], world)
      else:
        starts = np.cumsum([0] + list(depths[:-1]))
        parts = list(zip(starts.tolist(), depths))
      z0, zc = parts[rank]
      local = torch.from_numpy(vol[z0:z0 + zc].copy())
      out = ed.slab_transform(local, an, bb, sqrt=sqrt, signed=signed, passes=OraclePasses())
      queue.put((idx, rank, z0, out.numpy()))
  finally:
    dist.destroy_process_group()


CASES = [
  ((12, 9, 7), "iid", (1.0, 1.0, 1.0), False, False, False, None),
  ((12, 9, 7), "iid", (3.0, 2.0, 1.0), True, True, False, None),
  ((11, 5, 6), "blocks", (1.0, 2.0, 3.0), False, False, True, None),
  ((10, 4, 9), "ones", (1.0, 1.0, 1.0), False, False, False, None),       # inf-rich, long z runs
  ((10, 4, 9), "ones", (2.0, 1.0, 1.0), True, True, True, None),
  ((7, 3, 4), "blocks", (1.0, 1.0, 1.0), True, False, False, (7, 0)),     # an empty slab
  ((7, 3, 4), "blocks", (1.0, 1.0, 1.0), False, True, True, (0, 3, 4)),   # an empty first slab
]


@pytest.mark.parametrize("world", [2, 3])
def test_slab_split_matches_single_volume(world):
  sys.path.insert(0, ROOT)
  from oracle import oracle
  ctx = mp.get_context("spawn")
  queue = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
  for p in procs:
    p.start()
  active = [i for i, c in enumerate(CASES) if c[6] is None or len(c[6]) == world]
  results = [queue.get(timeout=300) for _ in range(world * len(active))]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert len(active) >= 6
  for idx in active:
    case = CASES[idx]
    shape, kind, an, bb, sqrt, signed, _ = case
    got = np.zeros(shape, dtype=np.float32)
    for i, rank, z0, arr in results:
      if i == idx:
        got[z0:z0 + arr.shape[0]] = arr
    fn = {(False, False): oracle.edtsq, (True, False): oracle.edt,
          (False, True): oracle.sdfsq, (True, True): oracle.sdf}[(sqrt, signed)]
    want = fn(_volume(case), anisotropy=an, black_border=bb)    # the same volume, not distributed
    assert np.array_equal(got, want, equal_nan=True), (world, idx)


def test_split_extent():
  sys.path.insert(0, ROOT)
  import edt_b200.distributed as ed
  assert ed.split_extent(10, 3) == [(0, 4), (4, 3), (7, 3)]
  assert ed.split_extent(2, 4) == [(0, 1), (1, 1), (2, 0), (2, 0)]
  assert sum(c for _, c in ed.split_extent(4096, 8)) == 4096


# ---------------------------------------------------------------------------------------
# the real thing: NCCL over 2 GPUs (skipped unless the box has at least two)
# ---------------------------------------------------------------------------------------

def _nccl_worker(rank, world, port, queue):
  sys.path.insert(0, ROOT)
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    import edt_b200
    import edt_b200.distributed as ed
    rng = np.random.default_rng(99)
    small = rng.integers(0, 5, (13, 6, 7))
    vol = np.repeat(np.repeat(np.repeat(small, 11, 0), 9, 1), 10, 2).astype(np.int32)   # 143 x 54 x 70
    vol[60:90, 10:30, 5:50] = rng.integers(0, 3, (30, 20, 45))
    for (bb, sqrt, signed, an) in ((False, False, False, (1.0, 1.0, 1.0)), (True, True, True, (3.0, 1.0, 2.0))):
      parts = ed.split_extent(vol.shape[0], world)
      z0, zc = parts[rank]
      local = torch.from_numpy(vol[z0:z0 + zc].copy()).cuda()
      out = ed.slab_transform(local, an, bb, sqrt=sqrt, signed=signed)
      whole = edt_b200.edt_cuda(torch.from_numpy(vol).cuda(), an, bb, sqrt=sqrt, signed=signed)
      torch.cuda.synchronize()
      queue.put((rank, bool(torch.equal(out, whole[z0:z0 + zc]))))
  finally:
    dist.destroy_process_group()


@pytest.mark.gpu
def test_slab_split_nccl_two_gpus():
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs")
  ctx = mp.get_context("spawn")
  queue = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, queue)) for r in range(2)]
  for p in procs:
    p.start()
  results = [queue.get(timeout=300) for _ in range(4)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert all(ok for _, ok in results), results
