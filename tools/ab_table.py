#!/usr/bin/env python
"""Tabulate the output of tools/ab.sh / tools/ab_env.sh: ms per (library, workload), both rounds."""
import re
import sys

for path in sys.argv[1:]:
  print(path)
  cur, res = None, {}
  for line in open(path):
    if line.startswith("=="):
      cur = re.sub(r" \(round \d\)", "", line.strip("= \n"))
      continue
    m = re.search(r'"workload": "(\w+)".*"ms": ([\d.]+)', line)
    if m:
      res.setdefault(cur, {}).setdefault(m.group(1), []).append(float(m.group(2)))
  wls = list(next(iter(res.values())).keys())
  print("%-44s" % "", " ".join("%12s" % w for w in wls))
  for k, v in res.items():
    print("%-44s" % k, " ".join("/".join("%5.3f" % t for t in v[w]).rjust(12) for w in wls))
