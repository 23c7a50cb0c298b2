"""Host-side multi-GPU logic on CPU: world_size-2 `gloo` process groups.

* batch enhancement shards by utterance / chain with no data-path collective: the sharded run (oracle injected as the
  numeric engine, since there is no GPU here) returns, on rank 0, exactly what one process returns, in corpus order;
* basis training shards frames and all-reduces [G | sum(H,2) | div] once per iteration: the sharded iteration equals
  the single-process oracle iteration (sparse_nmf.m:186-286) to float64 rounding.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _corpus():
    wavs = np.load(GOLDEN / "wavs.npz")
    m = wavs["M04_in"]
    rs = np.random.RandomState(5)
    pcms = [m[:160 * 30], m[9000:9000 + 160 * 12 + 7], np.zeros(0, np.int16), m[20000:20000 + 160 * 21],
            (rs.randn(160 * 9) * 2000).astype(np.int16)]
    ads = [rs.rand(50, 100) for _ in pcms]
    return pcms, ads


def _oracle_engine(ads, chain_mode):
    from oracle import snmf_oracle as O
    bases = np.load(GOLDEN / "bases.npz")
    rng = np.load(GOLDEN / "rng_seed1.npz")
    p = O.default_params()

    def engine(pcms_local, idx_local, chain_local):
        if not chain_mode:
            return [O.enhance_utterance(x, p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=rng["h_init"],
                                        Ad_blk_init=ads[i])[0] for x, i in zip(pcms_local, idx_local)]
        outs, k = [], 0
        while k < len(idx_local):            # consecutive utterances of one chain run as a chain
            j = k
            while j + 1 < len(idx_local) and chain_local[j + 1] == chain_local[k] and chain_local[k] >= 0:
                j += 1
            o, _ = O.enhance_chain(pcms_local[k:j + 1], p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=rng["h_init"],
                                   Ad_blk_inits=[ads[i] for i in idx_local[k:j + 1]])
            outs.extend(o)
            k = j + 1
        return outs
    return engine


def _enhance_worker(rank, world, port, chain_mode, q):
    import torch.distributed as dist
    from se_snmf_nat_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pcms, ads = _corpus()
        chain = [0, 1, 0, 1, -1] if chain_mode else None
        outs = sharding.enhance_corpus(pcms, _oracle_engine(ads, chain_mode), rank=rank, world=world, chain_id=chain)
        if rank == 0:
            q.put([o.tolist() for o in outs])
        else:
            assert outs is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chain_mode", [False, True], ids=["filewise", "chains"])
def test_sharded_enhancement_equals_single_process(chain_mode):
    import torch.multiprocessing as mp
    from se_snmf_nat_b200 import sharding
    pcms, ads = _corpus()
    chain = [0, 1, 0, 1, -1] if chain_mode else None
    single = sharding.enhance_corpus(pcms, _oracle_engine(ads, chain_mode), rank=0, world=1, chain_id=chain)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_enhance_worker, args=(r, 2, port, chain_mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(got) == len(single) == len(pcms)
    for a, b, x in zip(got, single, pcms):
        assert len(a) == (len(x) // 160 + 1) * 160
        assert np.array_equal(np.asarray(a, np.int16), b)


def test_shard_utterances_properties():
    from se_snmf_nat_b200 import sharding
    rs = np.random.RandomState(0)
    lens = (rs.uniform(3, 12, 1024) * 16000).astype(int)
    for world in (1, 2, 4, 8):
        sh = sharding.shard_utterances(lens, world)
        flat = sorted(i for s in sh for i in s)
        assert flat == list(range(1024))                       # a partition
        loads = [sum(lens[i] // 160 + 4 for i in s) for s in sh]
        assert max(loads) - min(loads) <= max(lens) // 160 + 4  # LPT: within one unit of balanced
    # chains stay whole, ordered, on one rank
    chain = [i // 4 for i in range(64)]
    sh = sharding.shard_utterances(lens[:64], 4, chain)
    for s in sh:
        for c in set(chain[i] for i in s):
            members = [i for i in s if chain[i] == c]
            assert members == [i for i in range(64) if chain[i] == c]
    assert sharding.shard_frames(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.shard_frames(10_000_000, 8)[-1] == (8_750_000, 10_000_000)


def _train_worker(rank, world, port, q):
    import torch.distributed as dist
    from se_snmf_nat_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, w, h, sparsity, iters = _train_problem()
        F, K = w.shape
        t0, t1 = sharding.shard_frames(V.shape[1], world)[rank]
        v = np.maximum(V[:, t0:t1], 1e-9)
        h = h[:, t0:t1].copy()
        costs = []
        for _ in range(iters):
            lam = np.maximum(w @ h, 1e-9)
            h = h * (w.T @ (v / lam)) / np.maximum(w.sum(0)[:, None] + sparsity, 1e-9)      # sparse_nmf.m:191-195
            lam = np.maximum(w @ h, 1e-9)
            buf = sharding.pack_accumulators((v / lam) @ h.T, h.sum(1), 0.0)
            G, hs, _ = sharding.unpack_accumulators(sharding.allreduce_sum(buf), F, K)
            w = sharding.w_update_from_accumulators(w, G, hs)
            lam = np.maximum(w @ h, 1e-9)
            div = float((v * np.log(v / lam) - v + lam).sum())
            tot = sharding.allreduce_sum(np.array([div, h.sum()]))
            costs.append(tot[0] + sparsity * tot[1])                                     # :250,261
        q.put((rank, w.tolist(), h.tolist(), costs))
    finally:
        dist.destroy_process_group()


def _train_problem():
    rs = np.random.RandomState(3)
    F, K, T = 65, 8, 101
    Wt = np.abs(rs.randn(F, K))
    V = Wt @ rs.gamma(0.3, 1.0, (K, T)) + 1e-9
    w0 = V[:, rs.choice(T, K, replace=False)].copy()
    w0 /= np.sqrt((w0 ** 2).sum(0))
    return V, w0, rs.rand(K, T), 5.0, 4


def test_frame_sharded_training_iteration_equals_oracle():
    import torch.multiprocessing as mp
    from oracle import snmf_oracle as O
    V, w0, h0, sparsity, iters = _train_problem()
    K = w0.shape[1]
    w_ref, h_ref, obj = O.sparse_nmf(V, init_w=w0, init_h=h0, max_iter=iters, sparsity=sparsity, conv_eps=0.0, cf="kl",
                                     w_update_ind=np.ones(K, bool), h_update_ind=np.ones(K, bool), cost_check=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w_a, w_b = np.array(res[0][1]), np.array(res[1][1])
    assert np.array_equal(w_a, w_b)                                   # replicated dictionary is bit-identical
    np.testing.assert_allclose(w_a, w_ref, rtol=1e-10, atol=1e-14)
    h = np.concatenate([np.array(res[0][2]), np.array(res[1][2])], axis=1)
    np.testing.assert_allclose(h, h_ref, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(res[0][3], obj["cost"], rtol=1e-10)
    assert res[0][3] == res[1][3]                                      # identical stop decisions on every rank
