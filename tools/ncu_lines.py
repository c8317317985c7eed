"""Per-source-line warp-sampling summary of an `ncu --set full --import-source on` report of the persistent kernel.

  python tools/ncu_lines.py gpurun_out/r1g_persistent.ncu-rep > profiles/r1g_ncu_top_lines.txt
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
addr2line, cur_file, cur_line, text = {}, None, None, {}
for r in csv.reader(cs.splitlines()):
  if not r:
    continue
  if r[0] == "File Path":
    cur_file = r[1].split("/")[-1]
    continue
  if r[0] in ("Function Name", "Line No"):
    continue
  if r[0] != "":
    cur_line = int(r[0])
    text[(cur_file, cur_line)] = r[1].strip()
  if len(r) > 2 and r[2].startswith("0x"):
    addr2line[r[2]] = (cur_file, cur_line)
sass = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = sass[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for r in sass[2:]:
  if len(r) < len(hdr):
    continue
  key = addr2line.get(r[0])
  if not key:
    continue
  a = agg[key]
  a["samples"] += int(r[idx["# Samples"]] or 0)
  a["inst"] += int(r[idx["Instructions Executed"]] or 0)
  for h in stalls:
    a[h] += int(r[idx[h]] or 0)
total = sum(a["samples"] for a in agg.values())
print(f"{rep}: {total} warp samples; top source lines (file:line, samples, share, warp instructions executed, main stall reasons)")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:40]:
  top = ", ".join(f"{h[6:]} {a[h]}" for h in sorted(stalls, key=lambda h: -a[h])[:3] if a[h])
  print(f"{key[0]}:{key[1]:<5d} {a['samples']:6d} {100 * a['samples'] / total:5.1f}%  inst {a['inst']:10d}  [{top}]  {text.get(key, '')[:90]}")
tot_st = collections.Counter()
for a in agg.values():
  for h in stalls:
    tot_st[h] += a[h]
print("all lines, by stall reason:", ", ".join(f"{h[6:]} {100 * v / max(1, sum(tot_st.values())):.1f}%" for h, v in tot_st.most_common(8)))
