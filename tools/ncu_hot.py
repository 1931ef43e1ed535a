"""Top stall sites of an .ncu-rep source page: python tools/ncu_hot.py <rep> [n]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
si, ii = h.index("# Samples"), h.index("Instructions Executed")
body = rows[2:]
tot = sum(int(r[si]) for r in body)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(f"total samples {tot}")
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:n]
for i in sorted(idx):
    r = body[i]
    print(f"{i:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}%  exec={r[ii]:>8s}  {r[1].strip()[:110]}")
