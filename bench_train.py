"""Training leg of bench.py (`--workload train`): BASELINE.json configs[3], `run_basis_train` -> `sparse_nmf` with W and H
both updated (run_basis_train.m:84-88, sparse_nmf.m:186-286), F=513, K=256, frames sharded over the ranks, one NCCL
all-reduce of the F x K accumulators per iteration.

A "step" is one multiplicative-update iteration (H-update + W-update) over all resident frames.  The synthetic
spectrogram follows SURVEY.md 8(d) config 4: V = W* H* + 1e-9, W* = |N(0,1)| column-normalised, H* ~ Gamma(0.3),
exemplar init (K random frames of V), H init U(0,1) -- generated on the device in chunks straight into the library's
resident arrays (snmfnat_train_dev_ptr), outside the timed region.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

METRIC = "MU iters/s: sparse_nmf dictionary-training iterations per second (run_basis_train, F=513, K=256)"


def _cudart():
    for nm in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(nm)
        except OSError:
            continue
    import torch  # noqa: F401  (torch ships a runtime)
    import glob
    import os.path as op
    import nvidia.cuda_runtime as rt
    return C.CDLL(glob.glob(op.join(op.dirname(rt.__file__), "lib", "libcudart.so*"))[0])


def fill_shard(tr, F, K, T, rank, dev):
    """Fill V [T][ldv], H [T][Kp] and W_init [K][F] of `tr` in place (device generated, chunked)."""
    import torch
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    D2D = 3
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    Wt = torch.randn(F, K, generator=g, device=dev, dtype=torch.float32).abs_()
    Wt /= Wt.norm(dim=0, keepdim=True)
    ldv, Kp = tr.ldv, tr.Kp
    pV, pH, pW = tr.dev_ptr("V"), tr.dev_ptr("H"), tr.dev_ptr("W_init")
    gam = torch.distributions.Gamma(torch.tensor(0.3, device=dev), torch.tensor(1.0, device=dev))
    chunk = 1 << 17
    first = None
    for t0 in range(0, T, chunk):
        n = min(chunk, T - t0)
        torch.manual_seed(2 + 1000003 * rank + t0)
        Hs = gam.sample((n, K)).float()                        # [n][K] = H*'
        Vc = torch.zeros(n, ldv, device=dev, dtype=torch.float32)
        Vc[:, :F] = Hs @ Wt.t() + 1e-9
        if first is None:
            first = Vc[:, :F].clone()
        Hc = torch.zeros(n, Kp, device=dev, dtype=torch.float32)
        Hc[:, :K] = torch.rand(n, K, generator=g, device=dev)
        torch.cuda.synchronize()
        assert rt.cudaMemcpy(C.c_void_p(pV + t0 * ldv * 4), C.c_void_p(Vc.data_ptr()), n * ldv * 4, D2D) == 0
        assert rt.cudaMemcpy(C.c_void_p(pH + t0 * Kp * 4), C.c_void_p(Hc.data_ptr()), n * Kp * 4, D2D) == 0
    # exemplar init (run_basis_train.m:80-83): K frames of V; every rank must start from the SAME dictionary, so the
    # exemplars come from the deterministic head of rank 0's generator (seed 3 picks rows of the first chunk)
    torch.manual_seed(2)
    Hs0 = gam.sample((min(chunk, T), K)).float()
    V0 = Hs0 @ Wt.t() + 1e-9
    idx = torch.randperm(V0.shape[0], generator=torch.Generator().manual_seed(3))[:K].to(dev)
    W0 = V0[idx].contiguous()                                  # [K][F]
    torch.cuda.synchronize()
    assert rt.cudaMemcpy(C.c_void_p(pW), C.c_void_p(W0.data_ptr()), K * F * 4, D2D) == 0
    tr.commit_v()
    tr.reset()
    del Vc, Hc, Hs, first, V0, Hs0


def measure_tf32_peak(dev, n=8192, reps=10):
    """Dense TF32 rate of cuBLAS on this GPU (CUBLAS_COMPUTE_32F_FAST_TF32 through torch.matmul with allow_tf32): the
    roofline denominator of the training kernels, measured instead of assumed.  Best of `reps`, CUDA events."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev, dtype=torch.float32)
        b = torch.randn(n, n, device=dev, dtype=torch.float32)
        for _ in range(3):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, float(e0.elapsed_time(e1)))
        del a, b
        return 2.0 * n ** 3 / (best / 1e3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def run(args, rank, world, local_rank, ClockSampler, measured_peaks):
    line = measure(args, rank, world, local_rank, ClockSampler, measured_peaks)
    if line is not None:
        print(json.dumps(line), flush=True)


def measure(args, rank, world, local_rank, ClockSampler, measured_peaks):
    """Returns the JSON line (a dict) on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from se_snmf_nat_b200 import api

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    F, K = 513, int(args.train_k)
    T_total = int(args.train_frames)
    T = T_total // world if args.train_scaling == "strong" else T_total
    ctx = api.Context(local_rank)
    tr = api.Train(ctx, F, K, T, 5.0)
    if world > 1:
        uid = [api.Train.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        tr.attach_nccl(uid[0], rank, world)
    t_gen = time.perf_counter()
    fill_shard(tr, F, K, T, rank, dev)
    ctx.sync()
    t_gen = time.perf_counter() - t_gen
    stream = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    tr.iterate(args.warmup)
    ctx.sync()
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count
    barrier(); torch.cuda.synchronize(); ctx.sync()
    sampler.start()
    e0.record(stream)
    tr.iterate(args.steps)
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize(); barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - l0
    ms = float(e0.elapsed_time(e1))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    prof = tr.profile() if hasattr(tr, "profile") else None

    # e2e: the public call a user makes for a training run -- iterate with the objective reported every iteration
    # (sparse_nmf's [w,h,objective] return): one scalar D2H + stream sync per iteration, plus the final W download
    barrier(); torch.cuda.synchronize(); ctx.sync()
    t0 = time.perf_counter()
    out = tr.iterate(args.steps, want_cost=True)
    w = tr.get_w()
    ctx.sync(); barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    tf32_cublas = measure_tf32_peak(dev)
    if rank != 0:
        tr.close()
        return None
    pk = measured_peaks()
    frames_total = T * world
    flop_iter = 4.0 * 2.0 * F * K * frames_total                     # 4 products of 2*F*K*T per iteration
    tf32_peak = tf32_cublas
    ach = flop_iter / world / (ms_step / 1e3) / 1e12
    cost = out["cost"]
    line = {
        "metric": METRIC, "value": 1e3 / ms_step, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": args.train_scaling, "vs_baseline": None, "dtype": "tf32 operands, fp32 accumulate/state",
        "data": "synthetic",
        "config": {"workload": "run_basis_train sparse_nmf dictionary learning (BASELINE.json configs[3])",
                   "F": F, "K": K, "frames_total": frames_total, "frames_per_gpu": T, "sparsity": 5.0, "cf": "kl",
                   "l2": "per-iteration working set (V + H, %.1f GB per GPU) is far larger than the 126 MB L2" %
                         ((T * (tr.ldv + 2 * tr.Kp) * 4) / 1e9),
                   "parallelism": f"frame-sharded x{world}, one NCCL all-reduce of F*K+K floats per iteration"},
        "clocks": clocks,
        "frames_per_s": frames_total / (ms_step / 1e3),
        "e2e": {"value": 1e3 / ms_e2e, "unit": "iters/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 8 + F * K * 4 // args.steps,
                "api": "snmfnat_train_iterate with div/cost read back every iteration + snmfnat_train_get_w "
                       "(training data is loaded once and stays resident, as in the reference's sparse_nmf call)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                     "traffic": (578e6 + 481e6) * (T / 151552.0),
                     "traffic_source": "dram__bytes_read+write of hphase2_kernel (578 MB) and wphase2_kernel (481 MB) per "
                                       "151 552 frames from profiles/r01_train_{h,w}phase2_ncu_full.txt, scaled to the frames "
                                       "per GPU of this run; algorithmic minimum per iteration ~ (2 V + 4 H) = "
                                       "(2*F + 4*Kp) * 4 B per frame",
                     "kernel": "hphase2_kernel + wphase2_kernel (tcgen05 tf32)",
                     "work_per_iteration_per_gpu": flop_iter / world,
                     "peak_source": "measured in this run: cuBLAS TF32 GEMM 8192^3 (torch.matmul, allow_tf32), best of 10 "
                                    "(MEASURED_PEAKS.json holds no TF32 figure)",
                     "half_of_measured_bf16_sustained": pk.get("bf16_tflops_sustained", 0.0) / 2.0},
        "kernel_ms": prof,
        "objective": {"first": float(cost[0]), "last": float(cost[-1]), "max_rel_increase": float(np.max(np.diff(cost) / cost[:-1])),
                      "non_increasing": bool(np.all(np.diff(cost) <= 1e-5 * cost[:-1])),
                      "note": "fp32 accumulators: increases below 1e-5 relative are summation noise near convergence"},
        "w_checksum": float(np.abs(w).sum()), "workload_gen_s": t_gen,
    }
    tr.close()
    return line
