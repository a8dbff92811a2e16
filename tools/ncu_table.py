#!/usr/bin/env python
"""Markdown table from ncu summaries (tools/ncu_summary.py output):
  python tools/ncu_table.py "label=path_ncu.json[:launch]" ...   (peak from MEASURED_PEAKS.json)"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
print("| capture | kernel | ms | DRAM GB (r+w) | GB/s | of peak | issue active % | regs | lanes/inst | warp inst |")
print("|---|---|---|---|---|---|---|---|---|---|")
for spec in sys.argv[1:]:
  label, path = spec.split("=", 1)
  idx = 0
  if re.search(r":\d+$", path):
    path, idx = path.rsplit(":", 1)
    idx = int(idx)
  l = json.load(open(path))["launches"][idx]
  kern = re.sub(r"^void ", "", l["kernel"]).split("(")[0]
  ms = l["gpu__time_duration.sum"] * 1e3
  gb = l["dram_bytes_per_launch"] / 1e9
  print("| %s | `%s` | %.4f | %.3f | %.0f | %.2f | %.0f | %d | %.1f | %.3g |" % (
      label, kern, ms, gb, gb / ms * 1e3, gb / ms * 1e3 / peak,
      l["smsp__issue_active.avg.pct_of_peak_sustained_active"], l["launch__registers_per_thread"],
      l["smsp__thread_inst_executed_per_inst_executed.ratio"], l["smsp__inst_executed.sum"]))
