"""Per-source-line stall sampling summary from an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_source.py rep.ncu-rep [top_n] [launch_index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines = []
cur_file = None
hdr = None
total = 0
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr is None or not row[0].isdigit():
        continue
    d = dict(zip(hdr, row))
    try:
        s = int(d["# Samples"])
    except Exception:
        continue
    total += s
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    lines.append((s, cur_file, int(d["Line No"]), d["Source"].strip()[:90], int(d["Instructions Executed"] or 0), stalls))
lines.sort(key=lambda x: -x[0])
print(f"total samples {total}")
for s, f, ln, src, inst, st in lines[:top]:
    top_st = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{100*s/max(total,1):5.1f}%  {f}:{ln:<4d} inst={inst:<9d} {src}\n        {top_st}")
