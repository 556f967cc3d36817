"""Summarise an `ncu --page source --csv` dump (SASS view with --import-source on): instructions executed per source line
and per opcode class, so the CUDA-core work of a kernel can be attributed to its roles.
   ncu -i rep.ncu-rep --page source --csv --launch-skip N --launch-count 1 > src.csv; python tools/ncu_source_hot.py src.csv"""
import csv
import collections
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0
by_op = collections.Counter()
samples = collections.Counter()
recs = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(float(r[ix["Instructions Executed"]]))
    except ValueError:
        continue
    src = r[ix["Source"]]
    op = re.sub(r"^@!?U?P\d+\s+", "", src.strip()).split(" ")[0].split(".")[0]
    by_op[op] += n
    tot += n
    s = int(float(r[ix["# Samples"]] or 0))
    samples[op] += s
    recs.append((n, s, src.strip()))
print("total warp-instructions", tot)
for op, n in by_op.most_common(40):
    print(f"{op:12s} {n:12d} {100.0 * n / tot:5.1f}%   samples {samples[op]}")
if len(sys.argv) > 2:
    print("--- hottest instructions")
    for n, s, src in sorted(recs, reverse=True)[: int(sys.argv[2])]:
        print(n, s, src[:110])
