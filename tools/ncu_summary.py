#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into a small JSON: per profiled launch the duration,
DRAM bytes, throughput percentages, occupancy, instruction counts and top stall reasons.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01/name.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
  "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
  "sm__throughput.avg.pct_of_peak_sustained_elapsed",
  "sm__warps_active.avg.pct_of_peak_sustained_active",
  "smsp__issue_active.avg.pct_of_peak_sustained_active",
  "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
  "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
  "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
  "launch__shared_mem_per_block_dynamic",
  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
  "smsp__thread_inst_executed_per_inst_executed.ratio",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def main():
  rep, out = sys.argv[1], sys.argv[2]
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  launches = []
  for r in data:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEYS:
      if k in hdr:
        i = hdr.index(k)
        try:
          val = float(r[i].replace(",", ""))
        except ValueError:
          continue
        scale = UNIT_SCALE.get(units[i])
        if scale is not None and ("bytes" in k or "duration" in k):
          val *= scale
        d[k] = val
    stalls = {}
    for i, h in enumerate(hdr):
      if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
          stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
        except ValueError:
          pass
    d["top_stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
      d["dram_bytes_per_launch"] = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    launches.append(d)
  with open(out, "w") as fh:
    json.dump({"report": rep, "launches": launches}, fh, indent=1)
  for d in launches:
    print(d["kernel"][:50], "%.1f us" % (d.get("gpu__time_duration.sum", 0) * 1e6),
          "dram %.3f GB" % (d.get("dram_bytes_per_launch", 0) / 1e9),
          "issue %.0f%%" % d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
          "dram%% %.0f" % d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
          d["top_stalls_warps_per_issue"])


if __name__ == "__main__":
  main()
