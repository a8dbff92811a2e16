#!/usr/bin/env python
"""Device-resident timing of the transform over a matrix of synthetic workloads (GPU only).

  python tools/perf_matrix.py [--size 512] [--only cfg2,cfg2b] [--steps 5]

Prints one JSON line per workload: ms per transform, Mvoxels/s, algorithmic GB/s.
Labels are generated on the GPU with torch (seeded); C-contiguous (z, y, x) tensors, i.e.
x fastest, which is the layout the kernels work in.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edt_b200  # noqa: E402


def gen(name, n, dev):
  g = torch.Generator(device=dev)
  g.manual_seed(0)
  if name == "cfg2":        # iid labels 0..255
    return torch.randint(0, 256, (n, n, n), dtype=torch.int32, device=dev, generator=g), (1, 1, 1), False
  if name == "cfg2b":       # 32^3 constant blocks of random labels
    small = torch.randint(0, 256, (n // 32,) * 3, dtype=torch.int32, device=dev, generator=g)
    big = small.repeat_interleave(32, 0).repeat_interleave(32, 1).repeat_interleave(32, 2)
    return big.contiguous(), (1, 1, 1), False
  if name == "blocks8":     # 8^3 blocks
    small = torch.randint(0, 256, (n // 8,) * 3, dtype=torch.int32, device=dev, generator=g)
    big = small.repeat_interleave(8, 0).repeat_interleave(8, 1).repeat_interleave(8, 2)
    return big.contiguous(), (1, 1, 1), False
  if name == "cfg3":        # all ones uint8, anisotropic, black border
    return torch.ones((n, n, n), dtype=torch.uint8, device=dev), (6, 6, 30), True
  if name == "ones_nobb":   # all ones, no border: everything stays +inf
    return torch.ones((n, n, n), dtype=torch.uint8, device=dev), (1, 1, 1), False
  if name in ("balls", "voronoi"):
    ax = torch.arange(n, device=dev, dtype=torch.float32)
    z, y, x = ax.view(n, 1, 1), ax.view(1, n, 1), ax.view(1, 1, n)
    if name == "balls":     # 64 random balls, binary
      lab = torch.zeros((n, n, n), dtype=torch.uint8, device=dev)
      c = torch.rand((64, 3), device=dev, generator=g) * n
      r = (40 + 50 * torch.rand((64,), device=dev, generator=g)) * (n / 512.0)
      for k in range(64):
        lab |= (((z - c[k, 0]) ** 2 + (y - c[k, 1]) ** 2 + (x - c[k, 2]) ** 2) <= r[k] ** 2).to(torch.uint8)
      return lab, (1, 1, 1), False
    c = torch.rand((200, 3), device=dev, generator=g) * n      # voronoi cells of 200 seeds
    best = torch.full((n, n, n), float("inf"), device=dev)
    lab = torch.zeros((n, n, n), dtype=torch.int32, device=dev)
    for k in range(200):
      d = (z - c[k, 0]) ** 2 + (y - c[k, 1]) ** 2 + (x - c[k, 2]) ** 2
      m = d < best
      best = torch.where(m, d, best)
      lab = torch.where(m, torch.full_like(lab, k + 1), lab)
    return lab, (1, 1, 1), False
  raise ValueError(name)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--size", type=int, default=512)
  ap.add_argument("--only", default="cfg2,cfg2b,blocks8,cfg3,balls,voronoi,ones_nobb")
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--sqrt", action="store_true")
  args = ap.parse_args()
  dev = torch.device("cuda", 0)
  for name in args.only.split(","):
    lab, an, bb = gen(name, args.size, dev)
    out = torch.empty(lab.shape, dtype=torch.float32, device=dev)
    for _ in range(2):
      edt_b200.edt_cuda(lab, an, bb, sqrt=args.sqrt, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
      edt_b200.edt_cuda(lab, an, bb, sqrt=args.sqrt, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    nvox = lab.numel()
    L = lab.element_size()
    print(json.dumps({"workload": name, "size": args.size, "label_bytes": L, "ms": round(ms, 4),
                      "Mvox_s": round(nvox / ms / 1e3, 1),
                      "alg_GBps": round((3 * L + 20) * nvox / ms / 1e6, 1),
                      "slab_mb": os.environ.get("EDTB200_XY_SLAB_MB", "default"),
                      "finite_max": float(out[torch.isfinite(out)].max().item()) if torch.isfinite(out).any() else None}),
          flush=True)
    del lab, out
    torch.cuda.empty_cache()


if __name__ == "__main__":
  main()
