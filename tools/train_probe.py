"""Stage-by-stage check of the tensor-core training kernels against float64 NumPy (diagnosis aid).
    python tools/train_probe.py [F] [K] [T]"""
import os
import sys
from pathlib import Path

os.environ["SNMFNAT_TRAIN_DEBUG"] = "1"

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from se_snmf_nat_b200 import api  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 513
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
FLR = 1e-9


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def case(name, V, W0, H0, sparsity=5.0):
    tr = api.Train(api.get_context(0), F, K, T, sparsity)
    tr.set_data(V, W0, H0)
    v = np.maximum(V.astype(np.float64), FLR)
    wn = np.sqrt((W0.astype(np.float64) ** 2).sum(0))
    w = W0 / wn
    h = H0 * wn[:, None]
    print(f"[{name}] init: W rel {rel(tr.get_w(), w):.2e}  H rel {rel(tr.get_h(), h):.2e}")
    lam = np.maximum(w @ h, FLR)
    cost0 = float((v * np.log(v / lam) - v + lam).sum())
    h1 = h * (w.T @ (v / lam)) / np.maximum(w.sum(0)[:, None] + sparsity, FLR)
    out = tr.iterate(1, want_cost=True)
    hg = tr.get_h().astype(np.float64)
    print(f"[{name}] H-update: H' rel {rel(hg, h1):.2e}   max|H'| gpu {hg.max():.3e} ref {h1.max():.3e}")
    if rel(hg, h1) > 1e-2:
        bad = np.abs(hg - h1) > 1e-2 * np.abs(h1).max()
        print(f"    bad entries {bad.sum()} of {bad.size}; bad atoms {np.unique(np.where(bad)[0])[:20]} bad frames "
              f"{np.unique(np.where(bad)[1])[:20]}")
    # W phase checked against the GPU's own H'
    lam1 = np.maximum(w @ hg, FLR)
    G = (v / lam1) @ hg.T
    hs = hg.sum(1)
    gg, hsg = tr.get_acc()
    print(f"[{name}] W-phase: G rel {rel(gg, G):.2e} (rows<512 {rel(gg[:512], G[:512]):.2e}, last row "
          f"{rel(gg[-1], G[-1]):.2e})  hs rel {rel(hsg, hs):.2e}")
    if rel(gg, G) > 1e-2:
        bad = np.abs(gg - G) > 1e-2 * np.abs(G).max()
        print(f"    bad entries {bad.sum()} of {bad.size}; bad bins {np.unique(np.where(bad)[0])[:20]} bad atoms "
              f"{np.unique(np.where(bad)[1])[:20]}")
    dpw = np.maximum(hs[None, :] + (G * w).sum(0)[None, :] * w, FLR)
    dmw = G + (hs[None, :] * w).sum(0)[None, :] * w
    w1 = w * dmw / dpw
    w1 /= np.sqrt((w1 ** 2).sum(0))
    print(f"[{name}] W-update: W rel {rel(tr.get_w(), w1):.2e}")
    lam2 = np.maximum(w1 @ hg, FLR)
    div = float((v * np.log(v / lam2) - v + lam2).sum())
    print(f"[{name}] cost: div gpu {out['div'][0]:.6e} ref {div:.6e} rel {abs(out['div'][0]-div)/div:.2e};  "
          f"cost gpu {out['cost'][0]:.6e} ref {div + sparsity*hg.sum():.6e}   (initial-state div {cost0:.6e})")
    tr.close()


rs = np.random.RandomState(0)
# A: layout-insensitive constants
case("const", np.full((F, T), 2.0, np.float32), np.ones((F, K), np.float32), np.ones((K, T), np.float32))
# B: random
Wt = np.abs(rs.randn(F, K))
Ht = rs.gamma(0.3, 1.0, (K, T))
V = (Wt @ Ht + 1e-9).astype(np.float32)
case("random", V, V[:, rs.choice(T, K, replace=False)].copy(), rs.rand(K, T).astype(np.float32))
