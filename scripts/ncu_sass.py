"""Annotated SASS of one kernel from an ncu report: executed warp-instructions and stall samples per SASS line.
usage: python scripts/ncu_sass.py report.ncu-rep kernel-regex out.txt"""
import csv, io, subprocess, sys
rep, kre, out = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ia, isr, ism, ith = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
lines = []; tot = 0
from collections import Counter
c = Counter()
for i, r in enumerate(rows[2:]):
    try: n = int(r[ia]); s = int(r[ism])
    except (ValueError, IndexError): continue
    src = r[isr].strip(); tot += n
    op = src.split()[1] if src.startswith('@') else src.split()[0]
    c[op.split('.')[0]] += n
    lines.append(f"{i:5d} {n/1e6:8.3f}M {s:6d} thr={r[ith]:>5} {src[:100]}")
open(out, "w").write("\n".join(lines))
print("total warp-instructions", tot)
for k, v in c.most_common(28): print(f"{k:12s} {v/1e6:8.1f}M {v/tot*100:5.1f}%")
