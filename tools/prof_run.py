"""Small fixed workload for ncu: N utterances x S seconds, one run of the batch path.
    python tools/prof_run.py [n_utt] [seconds]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench_workload as W  # noqa: E402
from se_snmf_nat_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
pcms, ads, fx = W.make_batch(n, rank=3, max_seconds=sec)
p = api.default_p()
ctx = api.Context(0)
b = api.Batch(ctx, p, fx["B_x"], fx["B_d"], [len(x) for x in pcms], fx["h_init"], ads)
b.upload(pcms)
b.set_profile(True)
b.run()
ctx.sync()
print(b.stats())
print(b.profile())
