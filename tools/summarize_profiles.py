"""Turns the raw ncu outputs of a gpurun call (gpurun_out/) into the tracked summaries under profiles/.

  python tools/summarize_profiles.py <launches.csv> <report.ncu-rep> <tag>
"""
import collections
import csv
import re
import subprocess
import sys

launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]

rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, data = rows[hi], rows[hi + 1 :]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in data:
  if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
    continue
  name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("mtx::", "")
  agg.setdefault(name, []).append(float(r[mv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(f"profiles/{tag}_launches_summary.csv", "w") as f:
  f.write("kernel,launches,mean_us,total_ms,share_pct\n")
  for k, v in agg.items():
    f.write(f"{k},{len(v)},{sum(v)/len(v)/1e3:.2f},{sum(v)/1e6:.3f},{100*sum(v)/tot:.1f}\n")
  f.write(f"TOTAL,{sum(len(v) for v in agg.values())},,{tot/1e6:.3f},100.0\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, body = rr[0], rr[1], rr[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__cluster_size", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
idx = [(w, h.index(w)) for w in want if w in h]
with open(f"profiles/{tag}_ncu_summary.csv", "w") as f:
  f.write(",".join(f"{w} [{units[i]}]" if units[i] else w for w, i in idx) + "\n")
  for r in body:
    f.write(",".join('"' + r[i].replace("mtx::", "")[:70] + '"' if w == "Kernel Name" else r[i] for w, i in idx) + "\n")
print(open(f"profiles/{tag}_launches_summary.csv").read())
print(open(f"profiles/{tag}_ncu_summary.csv").read()[:3000])
