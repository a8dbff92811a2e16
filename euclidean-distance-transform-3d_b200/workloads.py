"""Synthetic label volumes of the benchmark matrix, generated on the GPU.

Used by bench.py (the `workloads` array), tools/perf_matrix.py and the GPU tests, so that all of
them time and check the same volumes.  Every generator is seeded and returns a C-contiguous
(z, y, x) torch CUDA tensor -- x fastest, the layout the kernels work in -- together with the
anisotropy per ARRAY axis, i.e. (w_z, w_y, w_x) as `edt_cuda` takes it, and the black_border flag
of the configuration it stands for.

  cfg2     BASELINE configs[1]: iid labels 0..255, uint32, (1,1,1)        (runs of length ~1)
  cfg2b    32^3 constant blocks of random labels, uint32                  (blocky segmentation)
  blocks8  8^3 blocks
  cfg3     BASELINE configs[2]: all ones, uint8, (6,6,30), black border   (one run per line)
  balls    64 random balls of radius 40..90 (scaled with n), binary uint8 (large smooth objects)
  voronoi  Voronoi cells of 200 random seeds, uint32                      (dense segmentation with
                                                                           long runs of varying height)
  ones_nobb  all ones without a border: every distance stays +inf
"""
import torch

NAMES = ("cfg2", "cfg2b", "blocks8", "cfg3", "balls", "voronoi", "ones_nobb")


def generate(name, n, device, nz=None):
  """-> (labels[nz, n, n], anisotropy (w_z, w_y, w_x), black_border).  nz defaults to n."""
  nz = n if nz is None else nz
  g = torch.Generator(device=device)
  g.manual_seed(0)
  if name == "cfg2":
    return torch.randint(0, 256, (nz, n, n), dtype=torch.int32, device=device, generator=g), (1, 1, 1), False
  if name in ("cfg2b", "blocks8"):
    k = 32 if name == "cfg2b" else 8
    small = torch.randint(0, 256, (-(-nz // k), -(-n // k), -(-n // k)), dtype=torch.int32, device=device, generator=g)
    big = small.repeat_interleave(k, 0).repeat_interleave(k, 1).repeat_interleave(k, 2)
    return big[:nz, :n, :n].contiguous(), (1, 1, 1), False
  if name == "cfg3":
    return torch.ones((nz, n, n), dtype=torch.uint8, device=device), (30, 6, 6), True
  if name == "ones_nobb":
    return torch.ones((nz, n, n), dtype=torch.uint8, device=device), (1, 1, 1), False
  if name in ("balls", "voronoi"):
    z = torch.arange(nz, device=device, dtype=torch.float32).view(nz, 1, 1)
    y = torch.arange(n, device=device, dtype=torch.float32).view(1, n, 1)
    x = torch.arange(n, device=device, dtype=torch.float32).view(1, 1, n)
    scale = torch.tensor([nz, n, n], device=device, dtype=torch.float32)
    if name == "balls":
      lab = torch.zeros((nz, n, n), dtype=torch.uint8, device=device)
      c = torch.rand((64, 3), device=device, generator=g) * scale
      r = (40 + 50 * torch.rand((64,), device=device, generator=g)) * (n / 512.0)
      for k in range(64):
        lab |= (((z - c[k, 0]) ** 2 + (y - c[k, 1]) ** 2 + (x - c[k, 2]) ** 2) <= r[k] ** 2).to(torch.uint8)
      return lab, (1, 1, 1), False
    c = torch.rand((200, 3), device=device, generator=g) * scale
    best = torch.full((nz, n, n), float("inf"), device=device)
    lab = torch.zeros((nz, n, n), dtype=torch.int32, device=device)
    for k in range(200):
      d = (z - c[k, 0]) ** 2 + (y - c[k, 1]) ** 2 + (x - c[k, 2]) ** 2
      m = d < best
      best = torch.where(m, d, best)
      lab = torch.where(m, torch.full_like(lab, k + 1), lab)
    return lab, (1, 1, 1), False
  raise ValueError("unknown workload %r" % (name,))
