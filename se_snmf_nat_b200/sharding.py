"""Multi-GPU partitioning of the two paths that shard.  Host logic only: no CUDA in this module, so the
world_size-2 `gloo` tests exercise it on CPU with the numeric work injected.

* Batch enhancement: the unit is an utterance under `filewise_run_IS16` semantics (filewise_run_IS16.m:24-43,83: fresh
  bases per file) and a target-directory CHAIN under `Do_MultiBatch` / `NTF_sep_event_RT` semantics (the adapted noise
  basis flows file -> file through B_D_u.mat, src/NTF_sep_event_RT.m:28-38,136-139, reset per target directory,
  Do_MultiBatch_IS16_20160324_CHiME4.m:193).  Units go to ranks longest-processing-time-first; there is NO data-path
  collective, only the gather of the finished PCM to rank 0 when the caller wants it there.
* Basis training: frames split contiguously over ranks; one all-reduce of [G (F x K) | sum(H,2) (K) | div] per
  iteration (sparse_nmf.m:214-222,250), then every rank applies the same W-update.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


# ------------------------------------------------------------------------------------------ batch enhancement
def enhancement_units(lengths: Sequence[int], chain_id: Optional[Sequence[int]] = None):
    """Group utterance indices into independent units: singletons, or chains (same chain_id, original order kept)."""
    n = len(lengths)
    if chain_id is None:
        return [[i] for i in range(n)]
    units, where = [], {}
    for i in range(n):
        c = int(chain_id[i])
        if c < 0:
            units.append([i])
        elif c in where:
            units[where[c]].append(i)
        else:
            where[c] = len(units)
            units.append([i])
    return units


def shard_utterances(lengths: Sequence[int], world: int, chain_id: Optional[Sequence[int]] = None,
                     hop: int = 160, flush_hops: int = 4) -> List[List[int]]:
    """Longest-processing-time-first assignment of units to `world` ranks.  The cost of a unit is its hop count
    (filewise_run_IS16.m:102-123: len // hop + delay + 1 hops per file).  Deterministic: ties go to the lower rank and
    the lower unit index.  Returns, per rank, the utterance indices in the order they must run (chains stay contiguous
    and ordered)."""
    assert world >= 1
    units = enhancement_units(lengths, chain_id)
    cost = [sum(int(lengths[i]) // hop + flush_hops for i in u) for u in units]
    order = sorted(range(len(units)), key=lambda j: (-cost[j], j))
    load = [0] * world
    per_rank: List[List[int]] = [[] for _ in range(world)]
    for j in order:
        r = min(range(world), key=lambda q: (load[q], q))
        load[r] += cost[j]
        per_rank[r].append(j)
    out = []
    for r in range(world):
        idx: List[int] = []
        for j in sorted(per_rank[r]):   # keep corpus order inside a rank (stable outputs, chains untouched)
            idx.extend(units[j])
        out.append(idx)
    return out


def enhance_corpus(pcms: Sequence[np.ndarray], engine: Callable[[List[np.ndarray], List[int], Optional[List[int]]], List[np.ndarray]],
                   *, rank: int = 0, world: int = 1, chain_id: Optional[Sequence[int]] = None, group=None,
                   gather_to: Optional[int] = 0):
    """Enhance a corpus on `world` ranks.  `engine(pcms_local, indices_local, chain_local)` enhances this rank's
    utterances (on the GPU in production: `api.enhance_batch`; the tests inject a CPU checker) and returns their int16
    outputs in the same order.  With `gather_to` = r the full list (corpus order) is returned on rank r and None
    elsewhere; with gather_to=None every rank returns {index: output} of its own shard."""
    lengths = [len(x) for x in pcms]
    shards = shard_utterances(lengths, world, chain_id)
    mine = shards[rank]
    chain_local = None if chain_id is None else [int(chain_id[i]) for i in mine]
    outs = engine([pcms[i] for i in mine], list(mine), chain_local) if mine else []
    assert len(outs) == len(mine)
    local = {i: np.asarray(o) for i, o in zip(mine, outs)}
    if gather_to is None or world == 1:
        if world == 1 and gather_to is not None:
            return [local[i] for i in range(len(pcms))]
        return local
    import torch.distributed as dist
    gathered = [None] * world if rank == gather_to else None
    dist.gather_object(local, gathered, dst=gather_to, group=group)
    if rank != gather_to:
        return None
    full = {}
    for d in gathered:
        full.update(d)
    assert sorted(full) == list(range(len(pcms))), "an utterance was lost or duplicated by the sharding"
    return [full[i] for i in range(len(pcms))]


# ------------------------------------------------------------------------------------------ basis training
def shard_frames(T: int, world: int):
    """Contiguous frame ranges [t0, t1) per rank; the first T % world ranks get one extra frame."""
    base, rem = divmod(int(T), world)
    out, t = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((t, t + n))
        t += n
    return out


def pack_accumulators(G: np.ndarray, hs: np.ndarray, div: float) -> np.ndarray:
    """One fused all-reduce buffer per iteration: [G (F*K, column-major) | sum(H,2) (K) | div]."""
    return np.concatenate([np.asarray(G, np.float64).ravel(order="F"), np.asarray(hs, np.float64).ravel(), [float(div)]])


def unpack_accumulators(buf: np.ndarray, F: int, K: int):
    G = buf[:F * K].reshape((F, K), order="F")
    return G, buf[F * K:F * K + K], float(buf[F * K + K])


def w_update_from_accumulators(w: np.ndarray, G: np.ndarray, hs: np.ndarray, flr: float = 1e-9) -> np.ndarray:
    """sparse_nmf.m:214-222,242 (KL, every column updated) from the all-reduced accumulators: every rank computes the
    same new dictionary.  G = (V ./ Lambda) * H', hs = sum(H, 2)."""
    dpw = hs[None, :] + w * np.sum(G * w, axis=0, keepdims=True)
    dpw = np.maximum(dpw, flr)
    dmw = G + w * np.sum(hs[None, :] * w, axis=0, keepdims=True)
    w = w * dmw / dpw
    return w / np.sqrt(np.sum(w ** 2, axis=0, keepdims=True))


def allreduce_sum(buf: np.ndarray, group=None) -> np.ndarray:
    """Sum a host float64 buffer over the ranks of `group` (gloo on CPU in the tests; the GPU path all-reduces the
    device buffer with NCCL inside libsnmfnat, snmfnat_train_attach_nccl)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(buf))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
