"""Dynamic opcode mix + top stall lines from an .ncu-rep source page: python tools/ncu_opmix.py rep points [ncu launch filters]"""
import collections
import csv
import subprocess
import sys

rep, pts = sys.argv[1], float(sys.argv[2])
extra = sys.argv[3:]          # e.g. -s 2 -c 1 to pick one launch of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ci, ei, si = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
lsb, wait = hdr.index("stall_long_sb"), hdr.index("stall_wait")
cnt = collections.Counter()
tot = 0
lines = []
for r in rows[2:]:
    if len(r) <= ei:
        continue
    try:
        n = int(r[ei])
    except ValueError:
        continue
    op = r[ci].strip().split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    cnt[o.split(".")[0]] += n
    tot += n
    lines.append((int(r[si]), r[ci].strip(), n, int(r[lsb]), int(r[wait])))
for o, n in cnt.most_common(22):
    print(f"{o:12s} {n:14d} {n * 32 / pts:7.2f} thread-instr/pt")
print("total", tot, f"{tot * 32 / pts:.2f} thread-instr/pt")
print("--- top sampled lines (samples, long_sb, wait)")
for s, src, n, a, b in sorted(lines, reverse=True)[:25]:
    print(f"{s:7d} lsb={a:6d} wait={b:6d} exec={n:10d}  {src}")
