"""Rank source lines of an .ncu-rep by executed instructions (and show their sample share).
    python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; rows = []; cur = None
def num(x):
    try: return int(x)
    except Exception: return 0
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr and row[0].isdigit():
        d = dict(zip(hdr, row))
        rows.append((num(d["Instructions Executed"]), num(d["# Samples"]), cur, int(row[0]), d["Source"].strip()[:100]))
tot = sum(r[0] for r in rows); ts = sum(r[1] for r in rows)
print("total inst", tot, "samples", ts)
for r in sorted(rows, reverse=True)[:top]:
    print(f"{100*r[0]/max(tot,1):5.1f}% inst {100*r[1]/max(ts,1):5.1f}% smp  {r[2]}:{r[3]}  {r[4]}")
