"""Small fixed training workload for ncu: F=513, K=256, T frames, a few MU iterations.
    python tools/train_prof_run.py [T] [iters]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from se_snmf_nat_b200 import api  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
F, K = 513, 256
rs = np.random.RandomState(0)
Wt = np.abs(rs.randn(F, K)).astype(np.float32)
Wt /= np.linalg.norm(Wt, axis=0)
Ht = rs.gamma(0.3, 1.0, (K, T)).astype(np.float32)
V = Wt @ Ht + 1e-9
tr = api.Train(api.get_context(0), F, K, T, 5.0)
tr.set_data(V, V[:, rs.choice(T, K, replace=False)], rs.rand(K, T).astype(np.float32))
out = tr.iterate(iters, want_cost=True)
print(out)
tr.close()
