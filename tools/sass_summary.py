"""Per-kernel SASS evidence of libsnmfnat.so: counts of the instructions that prove the Blackwell-native paths
(tcgen05 -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, FP64 tensor cores -> DMMA, st.async over
distributed shared memory -> STAS, mbarrier -> SYNCS, cp.async -> LDGSTS).   python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "se_snmf_nat_b200" / "libsnmfnat.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "STAS", "SYNCS", "LDGSTS", "DFMA", "LDS", "BAR", "UCGABAR"]
cur = None
cnt = collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "")
        cur = re.sub(r"\(.*", "", cur)[:70]
        cnt[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1).split(".")[0]
        cnt[cur]["_total"] += 1
        if op in KEYS:
            cnt[cur][op] += 1
print(f"{'kernel':72s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
for k, c in cnt.items():
    if c["_total"] < 50:
        continue
    print(f"{k:72s} {c['_total']:6d} " + " ".join(f"{c[x]:7d}" for x in KEYS))
