"""Synthetic CHiME-4-shaped workload of BASELINE.json configs[2] (SURVEY.md 8d "Config 3").

1024 utterances of 16 kHz int16 noisy speech per GPU: durations ~U[3 s, 12 s] rounded to hops, each
synthesised as V = B_x*H_x + g*B_d'*H_d with the shipped speech dictionary, a column-permuted / perturbed
noise dictionary per utterance, Gamma(0.5) activations under an on/off Markov envelope (first 0.3 s
noise only, like CHiME), SNR ~U[0,15] dB, random phase, ISTFT + overlap-add, scaled to an int16 peak of
about 20000.  Utterances 0-2 are crops/tiles of the three real recordings the reference ships so that
real data is always in the batch.  Everything is seeded: utterance u of rank r uses seed 1000*r + u.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "tests" / "golden"
FS, HOP, WIN, NFFT = 16000, 160, 640, 1024


def load_fixtures():
    b = np.load(GOLDEN / "bases.npz")
    w = np.load(GOLDEN / "wavs.npz")
    r = np.load(GOLDEN / "rng_seed1.npz")
    return dict(B_x=b["B_DFT_x"], B_d=b["B_DFT_d"], wavs={k: w[k] for k in w.files}, h_init=r["h_init"],
                Ad_blk=r["Ad_blk"])


def durations(n_utt: int, seed: int = 0) -> np.ndarray:
    """Samples per utterance: U[3 s, 12 s] rounded to whole hops."""
    rs = np.random.RandomState(seed)
    hops = np.round(rs.uniform(3.0, 12.0, size=n_utt) * FS / HOP).astype(np.int64)
    return hops * HOP


def _markov_envelope(rs, T, p_pause=0.4, mean_run=25):
    env = np.empty(T)
    t = 0
    on = False  # starts with a noise-only stretch
    lead = int(0.3 * FS / HOP)
    env[:lead] = 0.0
    t = lead
    while t < T:
        run = 1 + rs.geometric(1.0 / mean_run)
        on = rs.rand() > p_pause if not on else False
        env[t:t + run] = 1.0 if on else 0.0
        t += run
    return env


def synth_utterance(n_samples: int, seed: int, B_x: np.ndarray, B_d: np.ndarray) -> np.ndarray:
    rs = np.random.RandomState(seed)
    T = n_samples // HOP + 3
    F = B_x.shape[0]
    # speech: few atoms active at a time (duty 5 %), temporally smooth, under the on/off envelope
    act = (rs.rand(B_x.shape[1], T) < 0.05).astype(np.float64)
    ker = np.ones(8) / 8.0
    act = np.stack([np.convolve(r, ker, mode="same") for r in act])
    Hx = rs.gamma(0.5, 1.0, size=(B_x.shape[1], T)) * act * _markov_envelope(rs, T)[None, :]
    perm = rs.permutation(B_d.shape[1])
    Bd = B_d[:, perm] * np.exp(0.3 * rs.randn(1, B_d.shape[1])) * np.exp(0.1 * rs.randn(F, 1))
    Hd = rs.gamma(0.5, 1.0, size=(B_d.shape[1], T)) * (0.5 + 0.5 * rs.rand(1, T))
    X = B_x @ Hx
    D = Bd @ Hd
    snr_db = rs.uniform(0.0, 15.0)
    g = np.sqrt(X.sum() / max(D.sum(), 1e-30) / 10 ** (snr_db / 10))
    V = X + (g * g) * D + 1e-9                      # power spectrogram
    spec = np.sqrt(V) * np.exp(2j * np.pi * rs.rand(F, T))
    spec[:5] = 0.0
    frames = np.fft.irfft(spec.T, n=NFFT, axis=1)[:, :WIN]
    k = np.arange(WIN)
    win = np.sqrt(0.5 * (1 - np.cos(2 * np.pi * k / WIN)))
    frames *= win[None, :]
    sig = np.zeros(T * HOP + WIN)
    for i in range(WIN // HOP):                     # overlap-add, 4 interleaved strided adds
        seg = frames[:, i * HOP:(i + 1) * HOP].reshape(-1)
        sig[i * HOP:i * HOP + T * HOP] += seg
    sig = sig[WIN - HOP:WIN - HOP + n_samples]
    peak = np.abs(sig).max()
    sig = sig * (20000.0 / max(peak, 1e-30))
    return np.clip(np.round(sig), -32768, 32767).astype(np.int16)


def synth_utterance_torch(n_samples: int, seed: int, B_x, B_d, device):
    """Same recipe as synth_utterance, generated on the GPU with torch (plumbing only: this is input
    synthesis, not the measured path).  B_x / B_d are float64 torch tensors on `device`."""
    import torch
    rs = np.random.RandomState(seed)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    T = n_samples // HOP + 3
    F, Rx = B_x.shape
    Rd = B_d.shape[1]
    env = torch.from_numpy(_markov_envelope(rs, T)).to(device)
    act = (torch.rand(Rx, T, generator=gen, device=device, dtype=torch.float64) < 0.05).to(torch.float64)
    act = torch.nn.functional.avg_pool1d(act[None], 8, stride=1, padding=4, count_include_pad=True)[0][:, :T]
    Hx = torch._standard_gamma(torch.full((Rx, T), 0.5, device=device, dtype=torch.float64), generator=gen) * act * env
    perm = torch.from_numpy(rs.permutation(Rd)).to(device)
    Bd = B_d[:, perm] * torch.exp(0.3 * torch.randn(1, Rd, generator=gen, device=device, dtype=torch.float64)) \
        * torch.exp(0.1 * torch.randn(F, 1, generator=gen, device=device, dtype=torch.float64))
    Hd = torch._standard_gamma(torch.full((Rd, T), 0.5, device=device, dtype=torch.float64), generator=gen) \
        * (0.5 + 0.5 * torch.rand(1, T, generator=gen, device=device, dtype=torch.float64))
    X = B_x @ Hx
    D = Bd @ Hd
    snr_db = rs.uniform(0.0, 15.0)
    g2 = X.sum() / torch.clamp(D.sum(), min=1e-30) / 10 ** (snr_db / 10)
    V = X + g2 * D + 1e-9
    ph = 2 * np.pi * torch.rand(F, T, generator=gen, device=device, dtype=torch.float64)
    spec = torch.polar(torch.sqrt(V), ph)
    spec[:5] = 0
    frames = torch.fft.irfft(spec.T, n=NFFT, dim=1)[:, :WIN]
    k = torch.arange(WIN, device=device, dtype=torch.float64)
    frames = frames * torch.sqrt(0.5 * (1 - torch.cos(2 * np.pi * k / WIN)))
    sig = torch.zeros(T * HOP + WIN, device=device, dtype=torch.float64)
    for i in range(WIN // HOP):
        sig[i * HOP:i * HOP + T * HOP] += frames[:, i * HOP:(i + 1) * HOP].reshape(-1)
    sig = sig[WIN - HOP:WIN - HOP + n_samples]
    sig = sig * (20000.0 / torch.clamp(sig.abs().max(), min=1e-30))
    return torch.clamp(torch.round(sig), -32768, 32767).to(torch.int16).cpu().numpy()


def real_utterance(wav: np.ndarray, n_samples: int) -> np.ndarray:
    reps = int(np.ceil(n_samples / len(wav)))
    return np.tile(wav, reps)[:n_samples].astype(np.int16)


def make_batch(n_utt: int, rank: int = 0, fixtures=None, max_seconds: float | None = None, device=None):
    """Returns (list of int16 arrays, Ad_blk inits (n_utt,50,100), fixtures).  With a torch CUDA `device` the
    synthesis runs on the GPU (seconds instead of minutes for 1024 utterances)."""
    fx = fixtures or load_fixtures()
    tb = None
    if device is not None:
        import torch
        tb = (torch.from_numpy(fx["B_x"]).to(device), torch.from_numpy(fx["B_d"]).to(device))
    lens = durations(n_utt, seed=rank)
    if max_seconds is not None:
        lens = np.minimum(lens, int(max_seconds * FS) // HOP * HOP)
    real = [fx["wavs"]["M03_in"], fx["wavs"]["M04_in"], fx["wavs"]["LM_in"]]
    pcms = []
    for u in range(n_utt):
        if u < 3 and rank == 0:
            pcms.append(real_utterance(real[u], int(lens[u])))
        elif tb is not None:
            pcms.append(synth_utterance_torch(int(lens[u]), 1000 * rank + u, tb[0], tb[1], device))
        else:
            pcms.append(synth_utterance(int(lens[u]), 1000 * rank + u, fx["B_x"], fx["B_d"]))
    rs = np.random.RandomState(77 + rank)
    ads = rs.rand(n_utt, 50, 100)
    ads[0] = fx["Ad_blk"]
    return pcms, ads, fx
