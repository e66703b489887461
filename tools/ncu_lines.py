"""Developer tool: aggregate an ncu report's source page by CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0] != "":
        try:
            agg.append((cur, int(r[0]), r[1], int(r[6]), int(r[7]), int(r[8])))
        except ValueError:
            pass
ti, tt, ts = sum(a[4] for a in agg), sum(a[5] for a in agg), sum(a[3] for a in agg)
print(f"warp-inst {ti:.4g}  thread-inst {tt:.4g}  avg lanes/inst {tt / ti:.2f}  samples {ts}")
byf = collections.defaultdict(lambda: [0, 0, 0])
for a in agg:
    for i in range(3):
        byf[a[0]][i] += a[3 + i]
for f, v in byf.items():
    print(f"{f:28s} samples {100 * v[0] / ts:5.1f}%  warp-inst {100 * v[1] / ti:5.1f}%  thread-inst {100 * v[2] / tt:5.1f}%  "
          f"lanes/inst {v[2] / max(1, v[1]):4.1f}")
agg.sort(key=lambda a: -a[3])
for a in agg[:top]:
    print(f"{a[0][:20]:20s} {a[1]:4d} smp {100 * a[3] / ts:5.2f}% winst {100 * a[4] / ti:5.2f}% tinst {100 * a[5] / tt:5.2f}% "
          f"lanes {a[5] / max(1, a[4]):4.1f} | {a[2][:80]}")
