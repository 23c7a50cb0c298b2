"""GPU parity of the tensor-core dictionary-training path (snmfnat_train_*: run_basis_train.m:80-91 ->
sparse_nmf.m:186-286 with W and H both updated) against the float64 oracle, through the C ABI.

Tolerance (BASELINE.json north_star): W / H relative error <= 1e-3 (Frobenius), cost within 1e-3.  The kernels compute
with tf32 operands and fp32 accumulation, so unlike the float64 online path they are not expected near 1e-12."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def api():
    from se_snmf_nat_b200 import api as a
    return a


@pytest.fixture(scope="module")
def O():
    from oracle import snmf_oracle
    return snmf_oracle


def make_problem(F, K, T, seed=0, ktrue=None):
    """SURVEY.md 8(d) config-4 generator at test size: V = W* H* + 1e-9, exemplar init, uniform H init."""
    rs = np.random.RandomState(seed)
    ktrue = ktrue or K
    Wt = np.abs(rs.randn(F, ktrue))
    Wt /= np.linalg.norm(Wt, axis=0)
    Ht = rs.gamma(0.3, 1.0, (ktrue, T))
    V = Wt @ Ht + 1e-9
    idx = rs.choice(T, K, replace=False)
    return V, idx, rs.rand(K, T)


def oracle_run(O, V, idx, H0, iters, sparsity=5.0, conv_eps=0.0):
    K = H0.shape[0]
    return O.sparse_nmf(V.astype(np.float32).astype(np.float64), init_w=V[:, idx].astype(np.float32).astype(np.float64),
                        init_h=H0.astype(np.float32).astype(np.float64), max_iter=iters, sparsity=sparsity,
                        conv_eps=conv_eps, cf="kl", w_update_ind=np.ones(K, bool), h_update_ind=np.ones(K, bool),
                        cost_check=True)


@pytest.mark.parametrize("F,K,T,iters", [
    (513, 64, 1000, 6),     # ragged T (not a multiple of 128 or 32), tail bin, 32-row streamed tiles
    (513, 100, 640, 6),     # shipped rank (run_basis_train R=100): padded to 128 atoms
    (513, 256, 2048, 4),    # BASELINE config-4 rank: 16-row streamed tiles
    (64, 32, 300, 6),       # Mel-sized dictionary: one partial 128-bin chunk, no tail row, one column block
    (257, 40, 129, 5),      # F % 128 == 1 again at another size, T just over one tile
    (200, 48, 96, 5),       # F not a multiple of 16/32, T smaller than one tile
])
def test_iterations_match_oracle(api, O, F, K, T, iters):
    V, idx, H0 = make_problem(F, K, T, seed=F + K)
    w_ref, h_ref, obj = oracle_run(O, V, idx, H0, iters)
    tr = api.Train(api.get_context(0), F, K, T, 5.0)
    try:
        tr.set_data(V, V[:, idx], H0)
        out = tr.iterate(iters, want_cost=True)
        w, h = tr.get_w(), tr.get_h()
    finally:
        tr.close()
    assert np.isfinite(w).all() and np.isfinite(h).all()
    assert rel_err(w_ref, w) < TOL, rel_err(w_ref, w)
    assert rel_err(h_ref, h) < TOL, rel_err(h_ref, h)
    assert np.allclose(np.linalg.norm(w.astype(np.float64), axis=0), 1.0, atol=1e-5)      # sparse_nmf.m:242
    np.testing.assert_allclose(out["cost"], obj["cost"], rtol=TOL)
    np.testing.assert_allclose(out["div"], obj["div"], rtol=2 * TOL)


def test_cost_is_non_increasing_and_split_calls_agree(api, O):
    """KL multiplicative updates never increase the cost; n iterations in one call == the same n in two calls."""
    F, K, T = 513, 64, 1500
    V, idx, H0 = make_problem(F, K, T, seed=7)
    ctx = api.get_context(0)
    a = api.Train(ctx, F, K, T, 5.0)
    b = api.Train(ctx, F, K, T, 5.0)
    try:
        a.set_data(V, V[:, idx], H0)
        b.set_data(V, V[:, idx], H0)
        ca = a.iterate(8, want_cost=True)["cost"]
        b.iterate(3)
        b.iterate(5)
        assert np.all(np.diff(ca) <= 1e-6 * ca[:-1])
        assert np.array_equal(a.get_w(), b.get_w())          # deterministic: fixed-order reductions
        assert np.array_equal(a.get_h(), b.get_h())
    finally:
        a.close()
        b.close()


def test_run_stops_like_the_reference(api, O):
    """sparse_nmf.m:273-283: stop after iteration it > 1 when the relative cost change drops below conv_eps."""
    F, K, T = 513, 32, 900
    V, idx, H0 = make_problem(F, K, T, seed=3, ktrue=8)
    eps = 2e-2
    w_ref, h_ref, obj = oracle_run(O, V, idx, H0, 60, conv_eps=eps)
    assert 2 <= obj["iters"] < 60
    tr = api.Train(api.get_context(0), F, K, T, 5.0)
    try:
        tr.set_data(V, V[:, idx], H0)
        out = tr.run(60, eps)
        w, h = tr.get_w(), tr.get_h()
    finally:
        tr.close()
    assert out["iters"] == obj["iters"]
    np.testing.assert_allclose(out["cost"], obj["cost"], rtol=TOL)
    assert rel_err(w_ref, w) < TOL and rel_err(h_ref, h) < TOL


def test_basis_train_core_mirror(api, O):
    """Host mirror of run_basis_train.m:80-91,112-116 against the oracle's restatement of the same lines."""
    F, K, T = 513, 20, 700
    V, idx, H0 = make_problem(F, K, T, seed=11)
    V = V.astype(np.float32).astype(np.float64)
    H0 = H0.astype(np.float32).astype(np.float64)
    p = dict(max_iter=12, sparsity=5.0, conv_eps=1e-3, cf="kl", cost_check=1)
    B_ref, A_ref, obj = O.basis_train_core(V, K, idx, p, h_init=H0)
    B, A, out = api.basis_train_core(V, K, idx, p, h_init=H0)
    assert out["iters"] == obj["iters"]
    assert rel_err(B_ref, B) < TOL and rel_err(A_ref, A) < TOL
    assert B.min() >= 1e-9 and abs(np.linalg.norm(B[:, 0] - 1e-9) - 1.0) < 1e-5


def test_unsupported_configurations_fail_loudly(api):
    ctx = api.get_context(0)
    with pytest.raises(api.SnmfnatError) as e:
        api.Train(ctx, 513, 300, 1000, 5.0)        # rank beyond the tensor-memory layout
    assert e.value.code == -4
    with pytest.raises(api.SnmfnatError):
        api.basis_train_core(np.ones((8, 8)), 2, [0, 1], dict(cf="is", sparsity=0, max_iter=1, conv_eps=0), h_init=np.ones((2, 8)))


def test_long_run_drift_against_float64_oracle(api, O):
    """TF32 operands over many iterations at a size closer to production (T = 16 384 frames, 60 iterations): W, H and
    the cost must stay inside north_star's 1e-3 of the float64 oracle, not merely be non-increasing."""
    F, K, T, iters = 513, 64, 16384, 60
    V, idx, H0 = make_problem(F, K, T, seed=5, ktrue=48)
    w_ref, h_ref, obj = oracle_run(O, V, idx, H0, iters)
    tr = api.Train(api.get_context(0), F, K, T, 5.0)
    try:
        tr.set_data(V, V[:, idx], H0)
        out = tr.iterate(iters, want_cost=True)
        w, h = tr.get_w(), tr.get_h()
    finally:
        tr.close()
    ew, eh = rel_err(w_ref, w), rel_err(h_ref, h)
    print(f"60 iterations at T=16384: rel err W {ew:.2e}, H {eh:.2e}, cost {abs(out['cost'][-1] / obj['cost'][-1] - 1):.2e}")
    assert ew < TOL and eh < TOL, (ew, eh)
    np.testing.assert_allclose(out["cost"], obj["cost"], rtol=TOL)


_NCCL_RANK_SCRIPT = r"""
import os, sys, time, numpy as np
rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
os.environ["CUDA_VISIBLE_DEVICES"] = str(rank)
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from se_snmf_nat_b200 import api
from test_gpu_train import make_problem
F, K, T, iters = 513, 64, 2000, 6
V, idx, H0 = make_problem(F, K, T, seed=21)
uid_file = os.path.join(tmp, "uid.bin")
if rank == 0:
    uid = api.Train.nccl_unique_id()
    open(uid_file + ".tmp", "wb").write(uid); os.replace(uid_file + ".tmp", uid_file)
else:
    for _ in range(600):
        if os.path.exists(uid_file): break
        time.sleep(0.1)
    uid = open(uid_file, "rb").read()
lo, hi = rank * T // world, (rank + 1) * T // world          # contiguous frame shard (sharding.frame_ranges)
tr = api.Train(api.get_context(0), F, K, hi - lo, 5.0)
tr.attach_nccl(uid, rank, world)
tr.set_data(V[:, lo:hi], V[:, idx], H0[:, lo:hi])
out = tr.iterate(iters, want_cost=True)
np.savez(os.path.join(tmp, f"rank{{rank}}.npz"), w=tr.get_w(), h=tr.get_h(), cost=out["cost"])
tr.close()
print("rank", rank, "ok")
"""


def test_frame_sharded_training_over_nccl_equals_single_gpu(api, O, tmp_path):
    """frames split over two ranks (one process per GPU), ONE ncclAllReduce of the F x K + K
    accumulators per iteration (snmfnat_train_attach_nccl), every rank applies the same W update.  The result must equal
    the single-GPU run up to the fp32 summation order of the accumulators."""
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    F, K, T, iters = 513, 64, 2000, 6
    V, idx, H0 = make_problem(F, K, T, seed=21)
    one = api.Train(api.get_context(0), F, K, T, 5.0)
    try:
        one.set_data(V, V[:, idx], H0)
        c1 = one.iterate(iters, want_cost=True)["cost"]
        w1, h1 = one.get_w(), one.get_h()
    finally:
        one.close()
    script = _NCCL_RANK_SCRIPT.format(root=str(ROOT))
    procs = [subprocess.Popen([sys.executable, "-c", script, str(r), "2", str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["w"], r1["w"])                       # replicated W update: bit-identical on both ranks
    assert rel_err(w1, r0["w"]) < 1e-5, rel_err(w1, r0["w"])
    assert rel_err(h1, np.concatenate([r0["h"], r1["h"]], axis=1)) < 1e-5
    np.testing.assert_allclose(r0["cost"], c1, rtol=1e-5)
