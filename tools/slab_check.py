#!/usr/bin/env python
"""Slab-split parity check, one rank per GPU (run under torchrun / torch.distributed.run):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tools/slab_check.py [--depth 64] [--size 512] [--full]

A structured sx x sy x (depth*N) volume (32^3 blocks of labels with background; seeded, the same on
every rank) is transformed by the slab-split path -- edtsq, edt with a black border, and sdf -- and
every rank compares ITS slab bit for bit with the same rows of the single-GPU transform of the
whole volume.  --full uses BASELINE configs[4]'s geometry: 512 x 512 x 4096 over 8 GPUs (sdf; the
single-GPU reference transform of the 1 Gi-voxel volume runs on every rank's own GPU).
Exit status 0 = every rank equal; rank 0 prints one line.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edt_b200  # noqa: E402
import edt_b200.distributed as ed  # noqa: E402
from edt_b200 import workloads  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--depth", type=int, default=64)
  ap.add_argument("--size", type=int, default=512)
  ap.add_argument("--full", action="store_true")
  args = ap.parse_args()
  rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  dist.init_process_group("nccl", device_id=dev)
  depth = 4096 // world if args.full else args.depth
  n = 512 if args.full else args.size
  lab, _, _ = workloads.generate("cfg2b", n, dev, nz=depth * world)
  lab = (lab % 7).to(torch.int32)                       # label 0 = background
  mine = lab[rank * depth:(rank + 1) * depth].contiguous()
  cases = [("sdf", dict(sqrt=True, signed=True), (2.0, 1.0, 1.0), False)]
  if not args.full:
    cases += [("edtsq", dict(sqrt=False, signed=False), (1.0, 1.0, 1.0), False),
              ("edt", dict(sqrt=True, signed=False), (3.0, 1.0, 2.0), True)]
  ok, methods = True, []
  for name, kw, an, bb in cases:
    whole = edt_b200.edt_cuda(lab, an, bb, **kw)
    want = whole[rank * depth:(rank + 1) * depth].clone()
    del whole
    info = {}
    got = ed.slab_transform(mine, an, bb, info=info, depths=[depth] * world, **kw)
    same = bool(torch.equal(got, want))
    ok = ok and same
    methods.append("%s:%s%s" % (name, info.get("method"), "" if same else "(DIFFERS)"))
    del got, want
    torch.cuda.empty_cache()
  flag = torch.tensor([1 if ok else 0], device=dev)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  if rank == 0:
    print("slab_check world=%d volume=%dx%dx%d %s -> %s" % (world, n, n, depth * world, " ".join(methods),
                                                          "OK" if int(flag.item()) else "MISMATCH"), flush=True)
  dist.destroy_process_group()
  sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
  main()
