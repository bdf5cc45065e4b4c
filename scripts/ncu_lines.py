"""Aggregate an ncu report's source page per CUDA source line: stall samples and executed warp instructions.
usage: python scripts/ncu_lines.py report.ncu-rep [kernel-regex] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kre: cmd += ["--kernel-name", "regex:" + kre]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
cur_file = None; hdr = None; agg = {}; seen_kernel = 0
for row in csv.reader(io.StringIO(txt)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No": hdr = row; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr is None or len(row) <= max(si, ii): continue
    if row[0] == "":  # sass rows
        continue
    try:
        key = (cur_file, int(row[0]), row[1].strip()[:90])
        s = int(row[si]); i = int(row[ii])
    except ValueError:
        continue
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += i
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"total samples {ts}  total warp-instructions {ti}")
print("---- by stall samples")
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{v[0]/ts*100:5.1f}% smp {v[1]/ti*100:5.1f}% ins  {k[0]}:{k[1]}  {k[2]}")
print("---- by instructions")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{v[1]/ti*100:5.1f}% ins {v[0]/ts*100:5.1f}% smp  {k[0]}:{k[1]}  {k[2]}")
