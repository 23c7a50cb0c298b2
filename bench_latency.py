"""Latency path (BASELINE.json configs[1], SURVEY.md 8d "Config 2"): wav/LM_in.wav enhanced as ONE stream, hop by hop,
through the per-hop entry (`init_buff` + `bnmf_sep_event_RT_IS16`, i.e. snmfnat_stream_create / snmfnat_stream_step).

Every call is synchronous like the MATLAB function it replaces: 640 host samples in, 160 enhanced host samples out,
the state `g` stays on the device.  Reported: per-hop wall-clock latency (p50 / p99 / max, the 10 ms hop is the
real-time budget) and xRT of the whole file.  A strict frame recurrence does not shard: N GPUs run N replicas.

    python bench.py --workload latency [--steps K] [--warmup W]      (one step = the whole 17.7 s recording)
"""
from __future__ import annotations

import json
import time

import numpy as np

import bench_workload as W

FS, HOP, WIN = 16000, 160, 640


def run(args, rank, world, local_rank, ClockSampler):
    import torch
    from se_snmf_nat_b200 import api

    torch.cuda.set_device(local_rank)
    fx = W.load_fixtures()
    pcm = fx["wavs"]["LM_in"]
    Bx, Bd = fx["B_x"], fx["B_d"]
    p = api.default_p()
    n_full = len(pcm) // HOP
    n_hops = n_full + 4                      # filewise_run_IS16.m:86-169: four flush hops
    audio_s = len(pcm) / FS

    def one_pass():
        g = api.init_buff(Bx, Bd, Bx, Bd, p, Ad_blk_init=fx["Ad_blk"], device=local_rank)
        y = np.zeros(WIN)
        lat = np.empty(n_hops)
        out = np.empty(n_hops * HOP)
        h_it = w_it = 0
        t_all = time.perf_counter()
        for l in range(1, n_hops + 1):
            if l <= n_full:
                y = np.concatenate([y[HOP:], pcm[(l - 1) * HOP:l * HOP].astype(np.float64)])
            else:
                y = np.zeros(WIN)                      # flush hops, as in tests/test_gpu_l1.py and filewise_run_IS16.m
            t0 = time.perf_counter()
            _, _, xt, g = api.bnmf_sep_event_RT_IS16(y, l, g, p, h_init=fx["h_init"], nargout=1)
            lat[l - 1] = time.perf_counter() - t0
            out[(l - 1) * HOP:l * HOP] = xt[:HOP]
            st = g["stats"]
            h_it += int(st[0])
            w_it += int(st[3])
        wall = time.perf_counter() - t_all
        g.close()
        return lat, wall, h_it, w_it, float(np.abs(out).sum())

    ctx = api.get_context(local_rank)
    for _ in range(max(args.warmup_latency, 1)):
        one_pass()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count
    lats, walls = [], []
    for _ in range(args.steps):
        lat, wall, h_it, w_it, chk = one_pass()
        lats.append(lat)
        walls.append(wall)
    launches = ctx.launch_count - l0
    clocks = sampler.stop()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([max(walls)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
    else:
        worst = max(walls)
    if rank != 0:
        return
    lat = np.concatenate(lats) * 1e3
    ms_step = 1e3 * float(np.mean(walls))
    line = {
        "metric": "xRT: enhanced audio-seconds per second (IS16 SNMF-NAT, single stream, per-hop entry)",
        "value": world * audio_s / worst, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup_latency, 1), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "real (wav/LM_in.wav shipped by the reference, tests/golden/wavs.npz)",
        "config": {"workload": "LM_in.wav demo enhancement, single stream, per-hop online NMF (BASELINE.json configs[1], "
                               "latency path)", "audio_s": audio_s, "hops": n_hops,
                   "parallelism": f"replicas only x{world} (a frame recurrence does not shard)",
                   "l2": "single stream: its 1.4 MB of state stays in L2 by design; the timed region is host wall clock "
                         "per synchronous call"},
        "latency_ms": {"p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90)),
                       "p99": float(np.percentile(lat, 99)), "max": float(lat.max()), "mean": float(lat.mean()),
                       "budget_ms": 10.0, "hops_over_budget": int((lat > 10.0).sum()), "hops": int(lat.size)},
        "e2e": {"value": world * audio_s / worst, "unit": "audio-s/s", "h2d_bytes_per_step": n_hops * (WIN + 200) * 8,
                "d2h_bytes_per_step": n_hops * (WIN + 8) * 8,
                "api": "init_buff + bnmf_sep_event_RT_IS16 per hop (snmfnat_stream_step): host frame in, host frame out"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": None, "cpu_baseline": None,
        "stats": {"h_iters": h_it, "w_iters": w_it, "launches_per_hop": launches / max(args.steps * n_hops, 1)},
        "output_checksum": chk,
    }
    print(json.dumps(line), flush=True)
