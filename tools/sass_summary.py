#!/usr/bin/env python3
"""Per-kernel SASS evidence of the built library: counts of the Blackwell-native mnemonics (B200_PROFILING.md,
"What proves a Blackwell-native kernel") in every kernel of libmtx_b200.so.  Writes profiles/sass_summary.txt.

  python tools/sass_summary.py            # runs here: cuobjdump needs no GPU
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxtext_indextts2_b200 import _lib  # noqa: E402

MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "HMMA", "LDSM", "SYNCS", "UTCBAR", "REDG", "RED.", "ATOMG", "MUFU.EX2", "CCTL"]


def main():
  _lib.build()
  sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
  arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
  kernels = collections.OrderedDict()
  name = None
  for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
      name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
      kernels[name] = collections.Counter()
      continue
    if name is None:
      continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m:
      op = m.group(1)
      kernels[name]["instructions"] += 1
      for mn in MNEMONICS:
        if op.startswith(mn):
          kernels[name][mn] += 1
  out = [f"# cuobjdump -sass {os.path.relpath(_lib.LIB_PATH, ROOT)} : arch {', '.join(arch)}",
         "# tcgen05.mma -> UTCHMMA, TMA load/store/prefetch -> UTMALDG/UTMASTG/UTMAPF, bulk copy -> UBLKCP, tcgen05.ld/st -> LDTM/STTM,",
         "# mma.sync -> HMMA, ldmatrix -> LDSM, mbarrier -> SYNCS, tcgen05.commit -> UTCBAR", ""]
  cols = ["instructions"] + MNEMONICS
  out.append(f"{'kernel':58s} " + " ".join(f"{c[:9]:>9s}" for c in cols))
  for k, c in kernels.items():
    out.append(f"{k[:58]:58s} " + " ".join(f"{c.get(col, 0):9d}" for col in cols))
  path = os.path.join(ROOT, "profiles", "sass_summary.txt")
  with open(path, "w", encoding="utf-8") as f:
    f.write("\n".join(out) + "\n")
  print("\n".join(out))


if __name__ == "__main__":
  main()
