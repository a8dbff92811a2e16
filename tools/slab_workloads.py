#!/usr/bin/env python
"""Slab-split transform on labelled volumes of different structure (torchrun --nproc-per-node N
tools/slab_workloads.py): which Z-pass method `auto` ends up with, its time per step including
the verdict, and bit-equality with the always-exact transpose method.  One 512^3 int32 slab per rank."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edt_b200.distributed as ed  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev,
                        pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
S = 512
z0 = rank * S


def blocks(edge):
  g = torch.Generator(device="cuda"); g.manual_seed(5)
  n = S // edge
  small = torch.randint(0, 256, (n * world, n, n), dtype=torch.int32, device=dev, generator=g)
  mine = small[rank * n:(rank + 1) * n]
  return mine.repeat_interleave(edge, 0).repeat_interleave(edge, 1).repeat_interleave(edge, 2).contiguous()


def balls(count, rmin, rmax):
  g = torch.Generator(device="cpu"); g.manual_seed(9)
  c = torch.rand((count, 3), generator=g) * torch.tensor([S * world, S, S])
  r = rmin + torch.rand(count, generator=g) * (rmax - rmin)
  z = torch.arange(z0, z0 + S, device=dev, dtype=torch.float32)[:, None, None]
  y = torch.arange(S, device=dev, dtype=torch.float32)[None, :, None]
  x = torch.arange(S, device=dev, dtype=torch.float32)[None, None, :]
  out = torch.zeros((S, S, S), dtype=torch.int32, device=dev)
  for k in range(count):
    inside = (z - c[k, 0]) ** 2 + (y - c[k, 1]) ** 2 + (x - c[k, 2]) ** 2 < r[k] ** 2
    out[inside] = k + 1
  return out


def iid():
  g = torch.Generator(device="cuda"); g.manual_seed(rank)
  return torch.randint(0, 256, (S, S, S), dtype=torch.int32, device=dev, generator=g)


WORKLOADS = [("iid labels", iid), ("32^3 blocks", lambda: blocks(32)), ("64^3 blocks", lambda: blocks(64)),
             ("128^3 blocks", lambda: blocks(128)),
             ("balls r 10-28 (x%d00)" % (4 * world), lambda: balls(400 * world, 10, 28)),
             ("balls r 40-90 (x%d)" % (32 * world), lambda: balls(32 * world, 40, 90))]
passes = ed.CudaPasses(dev)
for name, make in WORKLOADS:
  lab = make()
  kw = dict(passes=passes, depths=[S] * world)
  exact = ed.slab_transform(lab, method="transpose", **kw)
  info = {}
  for _ in range(2):
    got = ed.slab_transform(lab, info=info, **kw)
  same = bool(torch.equal(got, exact))
  torch.cuda.synchronize(); dist.barrier()
  t0 = time.perf_counter()
  for _ in range(5):
    ed.slab_transform(lab, info=info, **kw)
  torch.cuda.synchronize()
  auto_ms = (time.perf_counter() - t0) / 5 * 1e3
  dist.barrier()
  t0 = time.perf_counter()
  for _ in range(3):
    ed.slab_transform(lab, method="transpose", **kw)
  torch.cuda.synchronize()
  tr_ms = (time.perf_counter() - t0) / 3 * 1e3
  flag = torch.tensor([1 if same else 0], device=dev)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  if rank == 0:
    print("%-26s auto -> %-9s (halo %3s) %7.2f ms/step (verdict read every step)   transpose %7.2f ms   equal to transpose: %s   max %.1f"
          % (name, info["method"], info.get("halo", "-"), auto_ms, tr_ms, bool(flag.item()), float(exact[torch.isfinite(exact)].max())), flush=True)
  del lab, exact, got
dist.destroy_process_group()
