"""Aggregates an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, func, data, seen_func = None, None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        continue
    if r[0] == "Line No":
        continue
    if r[0] and r[0].isdigit() and len(r) > 7:
        try:
            data.append((int(r[6]), int(r[7]), fname, int(r[0]), r[1].strip()[:100], func))
        except ValueError:
            pass
funcs = sorted({d[5] for d in data})
first = funcs[0] if funcs else None
data = [d for d in data if d[5] == first]
ts, ti = sum(d[0] for d in data), sum(d[1] for d in data)
print("function:", first)
print("total samples %d, warp instructions %d" % (ts, ti))
print("%7s %6s %9s %6s  %s" % ("samples", "%", "inst", "%", "line"))
for d in sorted(data, reverse=True)[:top]:
    print("%7d %6.1f %9d %6.1f  %s:%d  %s" % (d[0], 100 * d[0] / ts, d[1], 100 * d[1] / ti, d[2], d[3], d[4]))
