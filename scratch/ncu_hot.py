"""Hot CUDA source lines of one launch in an .ncu-rep (needs -lineinfo + --import-source on):
    python scratch/ncu_hot.py rep.ncu-rep [launch] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, h, lines = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        h = r
    elif h and r[0].isdigit():          # a CUDA source line with aggregated metrics
        lines.append((fname, r))
ia = h.index("# Samples")
iex = h.index("Instructions Executed")
stalls = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[ia]) for _, r in lines if r[ia].isdigit())
print(f"{tot} samples over {len(lines)} source lines")
for f, r in sorted(lines, key=lambda fr: -int(fr[1][ia]) if fr[1][ia].isdigit() else 0)[:top]:
    st = {c.replace("stall_", ""): int(r[i]) for i, c in stalls if r[i].isdigit() and int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f[:18]:18s}:{r[0]:>4s} {int(r[ia]):6d} {100 * int(r[ia]) / tot:5.1f}% ex={r[iex]:>9s}  {r[1].strip()[:80]:80s} {st}")
