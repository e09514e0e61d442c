"""Opcode histogram of a SASS address range of one kernel (static listing from cuobjdump).
usage: python tools/sass_hist.py listing.txt [lo_hex hi_hex] [--full]"""
import collections
import re
import sys

path = sys.argv[1]
args = [a for a in sys.argv[2:] if not a.startswith("--")]
full = "--full" in sys.argv
lo = int(args[0], 16) if args else 0
hi = int(args[1], 16) if len(args) > 1 else 1 << 60
pat = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);")
ops = collections.Counter()
n = 0
for line in open(path):
    m = pat.match(line)
    if not m:
        continue
    addr = int(m.group(1), 16)
    if addr < lo or addr > hi:
        continue
    toks = m.group(2).split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    if not full:
        op = op.split(".")[0]
    ops[op] += 1
    n += 1
print("instructions:", n)
for k, v in ops.most_common():
    print("%-28s %5d %5.1f%%" % (k, v, 100.0 * v / n))
