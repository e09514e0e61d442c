"""Executed-instruction histogram by SASS opcode from an ncu report (source page, SASS view).
usage: python tools/ncu_ops.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
i_src, i_exe, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, smp, tot = collections.Counter(), collections.Counter(), 0
for r in rows:
    if len(r) <= i_exe or not r[0].startswith("0x"):
        continue
    s = r[i_src].split()
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    n = int(r[i_exe])
    ops[op] += n
    smp[op] += int(r[i_smp])
    tot += n
print("warp instructions executed:", tot)
for k, v in ops.most_common(top):
    print(f"{k:12s} {v:10d} {100 * v / tot:5.1f}%  samples {smp[k]}")
