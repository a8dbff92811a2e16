#!/usr/bin/env python
"""Phase timeline of the slab-split step (torchrun --nproc-per-node N tools/slab_profile.py)."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edt_b200.distributed as ed

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                        pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
g = torch.Generator(device="cuda"); g.manual_seed(rank)
lab = torch.randint(0, 256, (512, 512, 512), dtype=torch.int32, device="cuda", generator=g)
passes = ed.CudaPasses(torch.device("cuda", local))
peer, why = ed.make_peer_halo(torch.device("cuda", local), 512, 512, torch.int32, 32)
if rank == 0: print("peer halo:", "ok" if peer is not None else why)
if os.environ.get("NO_PEER"): peer = None
for _ in range(3):
  ed.slab_transform(lab, passes=passes, peer_halo=peer)
torch.cuda.synchronize(); dist.barrier()
acc = {}
wall = []
for it in range(10):
  info = {"marks": []}
  t0 = time.perf_counter()
  ed.slab_transform(lab, passes=passes, info=info, depths=[512] * world, peer_halo=peer, defer_check=True)
  t1 = time.perf_counter()
  torch.cuda.synchronize()
  wall.append((t1 - t0) * 1e3)
  m = info["marks"]
  for (n0, e0), (n1, e1) in zip(m[:-1], m[1:]):
    acc.setdefault(n1, []).append(e0.elapsed_time(e1))
  acc.setdefault("TOTAL", []).append(m[0][1].elapsed_time(m[-1][1]))
if rank == 0:
  for k, v in acc.items():
    print("%-32s %.3f ms" % (k, sum(v) / len(v)))
  print("host wall per call %.3f ms, method %s" % (sum(wall) / len(wall), info["method"]))
dist.destroy_process_group()
