#!/usr/bin/env python
"""Attribute the instructions / stall samples of an ncu source-page export to regions of
edt_kernels.cuh (device functions and the `// ====` stage markers of the kernels).

  ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
  python tools/ncu_source_regions.py src.csv [euclidean-distance-transform-3d_b200/csrc/edt_kernels.cuh] [top_lines]
"""
import csv
import re
import sys


def num(x):
  try:
    return int(x)
  except ValueError:
    return 0


def main():
  path = sys.argv[1]
  src = sys.argv[2] if len(sys.argv) > 2 else "euclidean-distance-transform-3d_b200/csrc/edt_kernels.cuh"
  top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
  rows = list(csv.reader(open(path)))
  hdr = next(r for r in rows if r and r[0] == "Line No")
  ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
  agg = {}
  for r in rows:
    if len(r) <= ti or not r[0]:
      continue
    try:
      ln = int(r[0])
    except ValueError:
      continue
    a = agg.setdefault(ln, [0, 0, 0, r[1]])
    a[0] += num(r[ii]); a[1] += num(r[si]); a[2] += num(r[ti])
  tot = sum(a[0] for a in agg.values()) or 1
  ts = sum(a[1] for a in agg.values()) or 1
  marks = []
  for k, line in enumerate(open(src), 1):
    m = re.match(r"__device__ __forceinline__ \S+ (\w+)\(", line) or re.match(r"\s*// =+ (stage [^=]*?) =+", line)
    if m:
      marks.append((k, m.group(1).strip()))
    m = re.match(r"(first_axis\w*|later_axis\w*|face_\w+)\(", line)
    if m:
      marks.append((k, "kernel " + m.group(1)))
  marks.sort()
  print("total warp instructions %d, samples %d, avg active lanes %.1f" %
        (tot, ts, sum(a[2] for a in agg.values()) / tot))
  for idx, (start, name) in enumerate(marks):
    end = marks[idx + 1][0] if idx + 1 < len(marks) else 10 ** 9
    i = sum(a[0] for l, a in agg.items() if start <= l < end)
    s = sum(a[1] for l, a in agg.items() if start <= l < end)
    t = sum(a[2] for l, a in agg.items() if start <= l < end)
    if i or s:
      print("%-34s inst %5.1f%%  samples %5.1f%%  lanes %4.1f" % (name[:34], 100.0 * i / tot, 100.0 * s / ts, t / max(i, 1)))
  print("--- top lines by instructions")
  for l, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d inst %5.1f%% smp %5.1f%% lanes %4.1f  %s" % (l, 100.0 * a[0] / tot, 100.0 * a[1] / ts, a[2] / max(a[0], 1), a[3][:100]))
  print("--- top lines by samples")
  for l, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top // 2]:
    print("%5d inst %5.1f%% smp %5.1f%% lanes %4.1f  %s" % (l, 100.0 * a[0] / tot, 100.0 * a[1] / ts, a[2] / max(a[0], 1), a[3][:100]))


if __name__ == "__main__":
  main()
