"""CPU oracle for the SNMF-NAT enhancement hot path.  TEST INFRASTRUCTURE ONLY.

This file is a float64 NumPy restatement of the reference's MATLAB algorithm
(lordet01/SE_SNMF_NAT, files cited per function as ``file:line`` relative to
the reference root).  It exists so that the CUDA path can be checked against
it; it is never imported by the product package ``se_snmf_nat_b200`` -- only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs.

Parity pinning: the reference ships no unit-level
known-answer vectors.  The oracle is pinned end to end against the two
input->output wav pairs the reference ships (``wav/*_out_v3.9_18.wav``): exact
output length and a coarse SNR (the golden files were produced with MATLAB RNG
streams that cannot be regenerated here, and the pipeline has hard gates, so
they cannot pin tighter than ~15-25 dB).  At kernel granularity the reference
itself leaves parity unpinned; this oracle *is* the pin, and says so.

All random draws the reference takes from MATLAB's generators (``rand`` in
``sparse_nmf.m:121,126,134`` and ``init_buff.m:37-38``) are INPUTS here.

Conventions: arrays are ordinary NumPy (row-major storage) but indexed exactly
like the MATLAB matrices (``w`` is F x R, ``h`` is R x n, ...).
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

FLR = 1e-9  # sparse_nmf.m:166


# ----------------------------------------------------------------------------
# parameters (settings/initial_setting_SNMF_NAT.m)
# ----------------------------------------------------------------------------
def sqrt_hann_periodic(n: int) -> np.ndarray:
    """sqrt(hann(n,'periodic')) -- settings/initial_setting_SNMF_NAT.m:33,35."""
    k = np.arange(n, dtype=np.float64)
    return np.sqrt(0.5 * (1.0 - np.cos(2.0 * np.pi * k / n)))


def default_params() -> dict:
    """The shipped configuration, settings/initial_setting_SNMF_NAT.m:1-149."""
    p = dict(
        blk_len_sep=1, blk_hop_sep=1, Splice=0,
        fs=16000, wintime=0.040, hoptime=0.010, ch=1,
        framelength=640, frameshift=160, delay=3, fftlength=1024,
        F_order=64, overlapscale=0.5, pow=2.0,
        EVENT_NUM=1, EVENT_RANK=[1], NOISE_NUM=1, NOISE_RANK=[1],
        R_x=100, R_d=100, nonzerofloor=1e-9,
        adapt_train_N=1, init_N_len=15, R_a=50, m_a=100, overlap_m_a=0.01, Ar_up=1.0,
        blk_sparse=1, P_len_k=60, P_len_l=20, alpha_p=0.4, blk_gap=3,
        preemph=0.0, DCbin=5, DCbin_back=5,
        B_sep_mode='DFT', MelConv=1,
        cf='kl', sparsity=5.0, max_iter=100, conv_eps=1e-3, random_seed=1, cost_check=1,
        basis_update_N=0, basis_update_E=0,
        ENHANCE_METHOD='MMSE', alpha_eta=0.4, alpha_d=0.6, beta=1.0, beta_max=1000.0,
        sparsity_mdi=5.0, conv_eps_mdi=1e-5,
        cluster_buff=1, train_Exemplar=0, domain_DD=0,
    )
    p['win_STFT'] = sqrt_hann_periodic(p['framelength'])
    p['win_ISTFT'] = sqrt_hann_periodic(p['framelength'])
    return p


# ----------------------------------------------------------------------------
# sparse_nmf  (src/sparse_nmf.m:71-292)
# ----------------------------------------------------------------------------
def _beta_of(cf: Optional[str], beta: Optional[float]) -> float:
    """sparse_nmf.m:95-110."""
    if cf is None:
        cf = 'kl'
    if cf == 'is':
        return 0.0
    if cf == 'kl':
        return 1.0
    if cf == 'ed':
        return 2.0
    return 1.0 if beta is None else float(beta)


def _divergence(v, lam, b):
    """sparse_nmf.m:248-258."""
    if b == 1:
        return float(np.sum(v * np.log(v / lam) - v + lam))
    if b == 2:
        return float(np.sum((v - lam) ** 2))
    if b == 0:
        return float(np.sum(v / lam - np.log(v / lam) - 1))
    return float(np.sum(v ** b + (b - 1) * lam ** b - b * v * lam ** (b - 1)) / (b * (b - 1)))


def sparse_nmf(v, *, init_w=None, init_h=None, r=None, max_iter=100, sparsity=0.0,
               conv_eps=0.0, cf='kl', beta=None, w_update_ind=None, h_update_ind=None,
               cost_check=True, rand: Optional[Callable] = None, mask=None, soft_mask=False,
               return_lambda=False):
    """[w,h,objective] = sparse_nmf(v,p), src/sparse_nmf.m:71-292.

    ``rand(m, n)`` supplies what MATLAB's ``rand`` would return after
    ``rand('seed', p.random_seed)`` (sparse_nmf.m:112-114); it is only needed
    when ``init_w`` is absent/short or ``init_h`` is absent.

    ``mask`` switches on the per-iteration imputation of src/snmf_mdi.m:175,
    251-255 (binary ``Dm``) or src/snmf_mdi_Sm.m:175,252-260 (``soft_mask``);
    the caller (`snmf_mdi`) does the final gain match.
    """
    v = np.array(v, dtype=np.float64, copy=True)
    if v.ndim == 1:
        v = v[:, None]
    m, n = v.shape
    b = _beta_of(cf, beta)

    # sparse_nmf.m:116-131
    if init_w is None:
        if r is None:
            raise ValueError('Number of components or initialization must be given')
        w = np.array(rand(m, r), dtype=np.float64)
    else:
        w = np.array(init_w, dtype=np.float64, copy=True)
        ri = w.shape[1]
        if r is not None and ri < r:
            w = np.concatenate([w, np.asarray(rand(m, r - ri), dtype=np.float64)], axis=1)
        else:
            r = ri
    # sparse_nmf.m:133-140
    if init_h is None:
        h = np.array(rand(r, n), dtype=np.float64)
    elif isinstance(init_h, str) and init_h == 'ones':
        h = np.ones((r, n))
    else:
        h = np.array(init_h, dtype=np.float64, copy=True)
        if h.ndim == 1:
            h = h[:, None]
    # sparse_nmf.m:142-148
    w_ind = np.ones(r, bool) if w_update_ind is None else np.asarray(w_update_ind, bool).ravel()
    h_ind = np.ones(r, bool) if h_update_ind is None else np.asarray(h_update_ind, bool).ravel()
    # sparse_nmf.m:150-155
    sp = np.asarray(sparsity, dtype=np.float64)
    if sp.size == 1:
        sp = np.ones((r, n)) * float(sp.reshape(()))
    elif sp.ndim == 1 or sp.shape[1] == 1:
        sp = np.repeat(sp.reshape(r, 1), n, axis=1)
    # sparse_nmf.m:157-160
    wn = np.sqrt(np.sum(w ** 2, axis=0))
    w = w / wn
    h = h * wn[:, None]

    flr = FLR
    lam = np.maximum(w @ h, flr)                       # :167
    last_cost = np.inf
    if mask is None:
        v = np.maximum(v, flr)                         # :169
    else:
        mk = np.asarray(mask, dtype=np.float64)
        v = np.maximum(v * mk, flr)                    # snmf_mdi.m:175
    div_hist = np.zeros(max_iter)
    cost_hist = np.zeros(max_iter)
    update_h = int(h_ind.sum())
    update_w = int(w_ind.sum())
    its = max_iter
    for it in range(1, max_iter + 1):                  # :186
        if update_h > 0:                               # :189-208
            wi = w[:, h_ind]
            if b == 1:
                dph = np.maximum(np.sum(wi, axis=0)[:, None] + sp[h_ind, :], flr)
                dmh = wi.T @ (v / lam)
            elif b == 2:
                dph = np.maximum(wi.T @ lam + sp[h_ind, :], flr)
                dmh = wi.T @ v
            else:
                dph = np.maximum(wi.T @ lam ** (b - 1) + sp[h_ind, :], flr)
                dmh = wi.T @ (v * lam ** (b - 2))
            h[h_ind, :] = h[h_ind, :] * dmh / dph
            lam = np.maximum(w @ h, flr)
        if update_w > 0:                               # :212-244
            wi = w[:, w_ind]
            hi = h[w_ind, :]
            if b == 1:
                rh = (v / lam) @ hi.T
                hs = np.sum(hi, axis=1)
                dpw = np.maximum(hs[None, :] + np.sum(rh * wi, axis=0)[None, :] * wi, flr)
                dmw = rh + np.sum(hs[None, :] * wi, axis=0)[None, :] * wi
            elif b == 2:
                lh = lam @ hi.T
                vh = v @ hi.T
                dpw = np.maximum(lh + np.sum(vh * wi, axis=0)[None, :] * wi, flr)
                dmw = vh + np.sum(lh * wi, axis=0)[None, :] * wi
            else:
                lh = lam ** (b - 1) @ hi.T
                vh = (v * lam ** (b - 2)) @ hi.T
                dpw = np.maximum(lh + np.sum(vh * wi, axis=0)[None, :] * wi, flr)
                dmw = vh + np.sum(lh * wi, axis=0)[None, :] * wi
            w[:, w_ind] = wi * dmw / dpw
            w = w / np.sqrt(np.sum(w ** 2, axis=0))    # :242 (all columns)
            lam = np.maximum(w @ h, flr)
        if mask is not None:                           # snmf_mdi.m:251-255
            v_est = np.maximum(w @ h, flr)
            inv = (1.0 - mk) if soft_mask else (mk == 0).astype(np.float64)
            v = np.maximum(v * mk + v_est * inv, flr)
        div = _divergence(v, lam, b)                   # :248-258
        if cost_check:                                 # :260-285
            cost = div + float(np.sum(sp * h))
            div_hist[it - 1] = div
            cost_hist[it - 1] = cost
            if it > 1 and conv_eps > 0:
                e = abs(cost - last_cost) / last_cost
                if e < conv_eps:
                    its = it
                    break
            last_cost = cost
    objective = dict(div=div_hist[:its], cost=cost_hist[:its], iters=its)
    if mask is not None or return_lambda:
        objective['v'] = v
    return w, h, objective


def snmf_mdi(v, Dm, *, soft_mask=False, sparsity_mdi=0.0, conv_eps_mdi=0.0, **kw):
    """[v_MDI,h,objective] = snmf_mdi(v,Dm,p), src/snmf_mdi.m:80-306 (binary mask)
    and src/snmf_mdi_Sm.m (soft mask; same code with ``1-Sm`` for ``~Dm``)."""
    v0 = np.array(v, dtype=np.float64)
    if v0.ndim == 1:
        v0 = v0[:, None]
    mk = np.asarray(Dm, dtype=np.float64).reshape(v0.shape)
    w, h, obj = sparse_nmf(v0, sparsity=sparsity_mdi, conv_eps=conv_eps_mdi, mask=mk,
                           soft_mask=soft_mask, **kw)
    vv = obj.pop('v')
    v_est = np.maximum(w @ h, FLR)                                         # :298
    Nt = np.sum(vv * mk, axis=0) / np.maximum(np.sum(v_est * mk, axis=0), FLR)   # :301
    inv = (1.0 - mk) if soft_mask else (mk == 0).astype(np.float64)
    v_mdi = np.maximum(vv * mk + (Nt[None, :] * v_est) * inv, FLR)         # :302-303
    return v_mdi, h, obj


def dnmf_adapt(Y, D, B, *, R_x, R_d, rand, **kw):
    """B_a = DNMF_adapt(Y,D,B,p), src/DNMF_adapt.m:3-20."""
    r = R_x + R_d
    _, A_hat, _ = sparse_nmf(Y, init_w=B, w_update_ind=np.zeros(r, bool),
                             h_update_ind=np.ones(r, bool), rand=rand, **kw)
    B_a, _, _ = sparse_nmf(D, init_w=B[:, R_x:R_x + R_d], init_h=A_hat[R_x:R_x + R_d, :],
                           w_update_ind=np.ones(R_d, bool), h_update_ind=np.zeros(R_d, bool),
                           rand=rand, **kw)
    return B_a


def gist_ntf(p, B, S_mag, *, rand, A=None, variant_c=False):
    """[C,A] = GIST_NTF(p,B,S_mag), src/GIST_NTF.m:8-155 (``variant_c``: src/GIST_NTF_C.m, objective gated by
    p.cost_check).  S_mag Channel x N x M.  Only C is updated (C_UPDATE=1, A_UPDATE=0, :4-6)."""
    S = np.asarray(S_mag, dtype=np.float64)
    Ch, N, M = S.shape
    B = np.array(B, dtype=np.float64)
    K = B.shape[1]
    C = np.array(rand(Ch, K), dtype=np.float64)                       # :14
    A = np.ones((M, K)) if A is None else np.asarray(A, dtype=np.float64)   # :16
    flr = p['nonzerofloor']
    Bn = np.sqrt(np.sum(B ** 2, axis=0))                              # :27-29
    B = B / Bn
    C = C * Bn
    xhat = lambda: np.maximum(np.einsum('hk,nk,mk->hnm', C, B, A), flr)   # kr(A,B,C) summed over k, :40-42
    X = xhat()
    P = np.maximum(S / X, flr)
    div_l, cost_l, its, last = [], [], 0, None
    for i in range(1, p['max_iter'] + 1):
        PBA = np.maximum(np.einsum('hnm,nk,mk->hk', P, B, A), flr)    # :96-113
        OBA = np.maximum(np.einsum('nk,mk->k', B, A), flr)[None, :]
        C = np.maximum(C * PBA / (OBA + p['sparsity']), flr)          # :114-115
        X = xhat()
        P = np.maximum(S / X, flr)
        its = i
        if (not variant_c) or p['cost_check']:
            with np.errstate(divide='ignore', invalid='ignore'):
                div = float(np.sum(S * np.log(S / X) - S + X))        # :131
            cost = div + float(np.sum(p['sparsity'] * C))
            div_l.append(div)
            cost_l.append(cost)
            if i > 1 and p['conv_eps'] > 0 and abs(cost - last) / last < p['conv_eps']:
                break
            last = cost
    return C, A, dict(div=np.array(div_l), cost=np.array(cost_l), iters=its)


def run_basis_dnmf(x, d, B, p, *, rand, mel=False):
    """B_hat = run_basis_DNMF(x,d,B,p), run_basis_DNMF.m:3-55; with ``mel`` the Mel variant
    run_basis_DNMF_Mel.m:3-93 (features projected with mel_matrix(...)' before the three solves)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    d = np.asarray(d, dtype=np.float64).ravel()
    n = min(x.size, d.size)                                           # :5-9
    x, d = x[:n], d[:n]
    y = x + d
    melmat = mel_matrix(p['fs'], p['F_order'], p['fftlength'], 1, p['fs'] / 2).T if mel else None

    def feat(sig):
        S, _ = stft_fft(sig, p['framelength'], p['frameshift'], p['fftlength'], p['DCbin'], p['win_STFT'], p['preemph'])
        S = S[:, np.any(S != 0, axis=0)]                              # :14,23,32
        S = S ** p['pow'] + p['nonzerofloor']                         # :16,25,34 (Splice = 0: frame_splice is identity)
        return melmat @ S if mel else S
    X, D, Y = feat(x), feat(d), feat(y)
    R_x, R_d = p['R_x'], p['R_d']
    r = R_x + R_d
    kw = dict(max_iter=p['max_iter'], sparsity=p['sparsity'], conv_eps=p['conv_eps'], cf=p['cf'],
              cost_check=bool(p['cost_check']))
    _, A_hat, _ = sparse_nmf(Y, init_w=B, w_update_ind=np.zeros(r, bool), h_update_ind=np.ones(r, bool),
                             rand=rand, **kw)                         # :37-40
    B_x, _, _ = sparse_nmf(X, init_w=B[:, :R_x], init_h=A_hat[:R_x, :], w_update_ind=np.ones(R_x, bool),
                           h_update_ind=np.zeros(R_x, bool), **kw)    # :43-47
    B_d, _, _ = sparse_nmf(D, init_w=B[:, R_x:r], init_h=A_hat[R_x:r, :], w_update_ind=np.ones(R_d, bool),
                           h_update_ind=np.zeros(R_d, bool), **kw)    # :49-53
    return np.concatenate([B_x, B_d], axis=1)


# ----------------------------------------------------------------------------
# STFT / ISTFT
# ----------------------------------------------------------------------------
def _preemph(x, a):
    """filter([1 -a],1,x) with zero initial state (bnmf_sep_event_RT_IS16.m:67)."""
    y = np.array(x, dtype=np.float64, copy=True)
    y[1:] -= a * np.asarray(x, dtype=np.float64)[:-1]
    return y


def _deemph(x, a):
    """filter(1,[1 -a],x) with zero initial state (synth_ifft_buff.m:26)."""
    if a == 0:
        return x
    y = np.empty_like(x)
    acc = 0.0
    for i, xv in enumerate(x):
        acc = xv + a * acc
        y[i] = acc
    return y


def frame_stft(y, p):
    """Inline STFT of one 640-sample frame, bnmf_sep_event_RT_IS16.m:66-78.
    Returns (Ym, Yp) with Ym = |Y|^pow, DC bins zeroed, + nonzerofloor."""
    sz, fftlen = p['framelength'], p['fftlength']
    yy = _preemph(np.asarray(y, dtype=np.float64).ravel(), p['preemph'])
    yy = p['win_STFT'] * yy
    Y = np.fft.fft(np.concatenate([yy, np.zeros(fftlen - sz)]))
    half = fftlen // 2 + 1
    Yp = np.angle(Y[:half])
    Ym = np.abs(Y[:half]) ** p['pow']
    Ym[:p['DCbin']] = 0.0
    Ym = Ym + p['nonzerofloor']
    return Ym, Yp


def stft_fft(s, sz, shift, fftlen, DCbin, win, preemph):
    """[S_mag,S_phase] = stft_fft(...), src/stft_fft.m:15-37 (training STFT:
    magnitude not power, DC bins := 1e-6, loop bound drops the tail frames,
    output preallocated with floor(len/shift) columns)."""
    s = np.asarray(s, dtype=np.float64).ravel()
    frame_num = len(s) // shift
    half = fftlen // 2 + 1
    S_mag = np.zeros((half, frame_num))
    S_phase = np.zeros((half, frame_num))
    pos = 1  # 1-based like the reference
    i = 0
    while pos < len(s) - fftlen:
        fr = _preemph(s[pos - 1:pos - 1 + sz], preemph)
        fr = win * fr
        S = np.fft.fft(np.concatenate([fr, np.zeros(fftlen - sz)]))
        mag = np.abs(S[:half])
        S_phase[:, i] = np.angle(S[:half])
        mag[:DCbin] = 0.000001
        S_mag[:, i] = mag
        pos += shift
        i += 1
    return S_mag, S_phase


def synth_ifft_buff(TF_mag, TF_phase, sz, fftlen, win, preemph, DCbin_back, pow_):
    """s_buff = synth_ifft_buff(...), src/synth_ifft_buff.m:4-33."""
    TF_mag = np.array(TF_mag, dtype=np.float64, copy=True)
    TF_phase = np.asarray(TF_phase, dtype=np.float64)
    if TF_mag.ndim == 1:
        TF_mag = TF_mag[:, None]
        TF_phase = TF_phase.reshape(-1, 1)
    freq_num, frame_num = TF_mag.shape
    s_buff = np.zeros((sz, frame_num))
    TF_mag[:DCbin_back, :] = 0.0
    TF_mag = TF_mag ** (1.0 / pow_)
    for i in range(frame_num):
        if freq_num == fftlen:
            TF = TF_mag[:, i].astype(np.complex128)
        else:
            mag = np.concatenate([TF_mag[:, i], TF_mag[1:fftlen // 2, i][::-1]])
            ph = np.concatenate([TF_phase[:, i], -TF_phase[1:fftlen // 2, i][::-1]])
            TF = mag * np.exp(1j * ph)
        sp = np.real(np.fft.ifft(TF))[:sz]
        sp = sp * win
        s_buff[:, i] = _deemph(sp, preemph)
    return s_buff


# ----------------------------------------------------------------------------
# Mel helpers (src/mel_matrix.m:15-38) -- used by Mel mode and basis training
# ----------------------------------------------------------------------------
def _mround(x):
    """MATLAB round(): half away from zero."""
    x = np.asarray(x, dtype=np.float64)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def mel_matrix(fs, NbCh, Nfft, warp=1.0, fhigh=None):
    """M = mel_matrix(fs,NbCh,Nfft,warp,fhigh), src/mel_matrix.m:9-38.
    Returns the (Nfft/2+1) x NbCh triangular weight matrix (dense)."""
    if fhigh is None:
        fhigh = fs / 2
    LowMel = 2595 * np.log10(1 + 64 / 700)
    NyqMel = 2595 * np.log10(1 + fhigh / 700)
    StartMel = LowMel + np.arange(NbCh) / (NbCh + 1) * (NyqMel - LowMel)
    fCen = warp * 700 * (10 ** (StartMel / 2595) - 1)
    StartBin = _mround(Nfft / fs * fCen).astype(int) + 1
    EndMel = LowMel + np.arange(2, NbCh + 2) / (NbCh + 1) * (NyqMel - LowMel)
    EndBin = _mround(warp * Nfft / fs * 700 * (10 ** (EndMel / 2595) - 1)).astype(int) + 1
    TotLen = EndBin - StartBin + 1
    LowLen = np.concatenate([StartBin[1:NbCh], [EndBin[NbCh - 2]]]) - StartBin + 1
    HiLen = TotLen - LowLen + 1
    rows = max(int(math.ceil(warp * Nfft / 2 + 1)), int(EndBin.max()))
    M = np.zeros((rows, NbCh))
    for k in range(NbCh):
        sb, ll, eb, hl = StartBin[k], LowLen[k], EndBin[k], HiLen[k]
        M[sb - 1:sb - 1 + ll, k] = np.arange(1, ll + 1) / ll
        M[eb - hl:eb, k] = np.arange(hl, 0, -1) / hl
    return M[:Nfft // 2 + 1, :]


# ----------------------------------------------------------------------------
# blk_sparse  (src/blk_sparse.m:3-36)
# ----------------------------------------------------------------------------
def blk_sparse(X, D, r_blk, l, p):
    """[Q,r_blk_out] = blk_sparse(X,D,r_blk,l,p), src/blk_sparse.m:3-36."""
    X = np.asarray(X, dtype=np.float64).ravel()
    D = np.asarray(D, dtype=np.float64).ravel()
    K = X.shape[0]
    gapN2 = int((p['blk_gap'] - 1) // 2)
    snr = X / np.maximum(D, p['nonzerofloor'])
    snr = snr / np.max(snr)
    r_out = np.concatenate([r_blk[:, 1:p['P_len_l']], snr[:, None]], axis=1)
    Q = np.concatenate([np.zeros(p['DCbin']), 0.1 * np.ones(K - p['DCbin'])])
    n = p['P_len_l'] * p['P_len_k']
    k2 = int(math.floor(p['P_len_k'] * 0.5))
    if l > p['P_len_l']:
        k = k2 + p['DCbin']                        # 1-based
        while k <= K - k2:
            b = r_out[k - k2:k + k2, :]            # rows k-k2+1 .. k+k2 (1-based)
            l1 = np.sum(b)
            l2 = np.sqrt(np.sum(b ** 2))
            P_tmp = (math.sqrt(n) - l1 / l2) / (math.sqrt(n) - 1)
            P_val = p['alpha_p'] * Q[k - 2] + (1 - p['alpha_p']) * P_tmp
            Q[k - 1 - gapN2:k] = P_val
            Q[k - 1:k + gapN2] = P_val
            k += p['blk_gap']
        Q[:p['P_len_k'] - 1] = Q[p['P_len_k'] + p['DCbin'] - 1]
    Q[:p['DCbin']] = 0.0
    return Q, r_out


# ----------------------------------------------------------------------------
# state + per-hop step
# ----------------------------------------------------------------------------
@dataclass
class State:
    """The struct ``g`` of src/init_buff.m:17-62 (fields used by the IS16 path)."""
    B_Mel_x: np.ndarray
    B_Mel_d: np.ndarray
    B_DFT_x: np.ndarray
    B_DFT_d: np.ndarray
    A_d: np.ndarray
    Ad_blk: np.ndarray
    lambda_d_blk: np.ndarray
    r_blk: np.ndarray
    lambda_dav: np.ndarray
    Xm_tilde: np.ndarray
    Ym: np.ndarray
    Yp: np.ndarray
    update_switch: int = 1
    blk_cnt: int = 1
    melmat: Optional[np.ndarray] = None
    # diagnostics of the last hop (not part of the reference struct)
    dbg: dict = field(default_factory=dict)


def init_buff(B_Mel_x, B_Mel_d, B_DFT_x, B_DFT_d, p, *, Ad_blk_init, A_d_init=None) -> State:
    """g = init_buff(B_Mel_x,B_Mel_d,B_DFT_x,B_DFT_d,p), src/init_buff.m:5-62.
    ``Ad_blk_init`` (R_a x m_a) and ``A_d_init`` (R_d x 1) replace the two
    ``rand`` draws at :37-38."""
    n2, _ = B_DFT_x.shape
    R_d = B_DFT_d.shape[1]
    if p['blk_len_sep'] != 1 or p['Splice'] != 0:
        raise ValueError('blk_len_sep>1 / Splice>0 are not supported by the IS16 frame function')
    Ad = np.array(Ad_blk_init, dtype=np.float64, copy=True)
    assert Ad.shape == (p['R_a'], p['m_a'])
    g = State(
        B_Mel_x=np.array(B_Mel_x, dtype=np.float64, copy=True),
        B_Mel_d=np.array(B_Mel_d, dtype=np.float64, copy=True),
        B_DFT_x=np.array(B_DFT_x, dtype=np.float64, copy=True),
        B_DFT_d=np.array(B_DFT_d, dtype=np.float64, copy=True),
        A_d=np.zeros(R_d) if A_d_init is None else np.array(A_d_init, dtype=np.float64).ravel(),
        Ad_blk=Ad,
        lambda_d_blk=np.zeros((n2, p['m_a'])),
        r_blk=np.zeros((n2, p['P_len_l'])),
        lambda_dav=np.zeros(n2),
        Xm_tilde=np.zeros(n2),
        Ym=np.zeros(n2) + p['nonzerofloor'],
        Yp=np.zeros(n2),
    )
    if p['B_sep_mode'] == 'Mel':
        g.melmat = mel_matrix(p['fs'], p['F_order'], p['fftlength'], 1, p['fs'] / 2).T  # :61
    return g


def bnmf_sep_event_RT_IS16(y, l, g: State, p, *, h_init, want_aux=False):
    """[x_hat_i,d_hat_i,x_tilde,g] = bnmf_sep_event_RT_IS16(y,l,g,p),
    src/bnmf_sep_event_RT_IS16.m:47-420, for blk_len_sep=1, Splice=0.

    ``h_init`` (R x 1) is the vector MATLAB's ``rand(r,1)`` yields right after
    ``rand('seed',p.random_seed)`` in the H-solve (sparse_nmf.m:112-114,134).
    Mutates and returns ``g``.  Returns (x_hat_i, d_hat_i, x_tilde, g); the
    first two are None unless ``want_aux``.
    """
    B_DFT_x, B_DFT_d = g.B_DFT_x, g.B_DFT_d
    B_Mel_x, B_Mel_d = g.B_Mel_x, g.B_Mel_d
    n1 = B_Mel_d.shape[0]
    n2, R_x = B_DFT_x.shape
    R_d = B_DFT_d.shape[1]
    r = R_x + R_d
    mel = p['B_sep_mode'] == 'Mel'
    flr = p['nonzerofloor']
    nmf_kw = dict(max_iter=p['max_iter'], sparsity=p['sparsity'], conv_eps=p['conv_eps'],
                  cf=p['cf'], cost_check=bool(p['cost_check']))

    Ym, Yp = frame_stft(y, p)                                   # :66-78

    # Mel projection :107-122
    if mel:
        Ym_Mel = g.melmat @ Ym
        vn = np.sqrt(np.sum(Ym_Mel ** 2))
        tn = np.sqrt(np.sum(Ym ** 2))
        Ym_Mel = (Ym_Mel / vn + 1e-9) * tn
        Y_sep = Ym_Mel
    else:
        Y_sep = Ym

    # 1) H-solve :124-154
    if p['basis_update_N']:
        w_ind = np.concatenate([np.zeros(R_x, bool), np.ones(R_d, bool)])
    elif p['basis_update_E']:
        w_ind = np.concatenate([np.ones(R_x, bool), np.zeros(R_d, bool)])
    else:
        w_ind = np.zeros(r, bool)
    B_Mel = np.concatenate([B_Mel_x, B_Mel_d], axis=1)
    B_DFT = np.concatenate([B_DFT_x, B_DFT_d], axis=1)
    _, A, obj_h = sparse_nmf(Y_sep, init_w=B_Mel if mel else B_DFT, init_h=np.asarray(h_init).reshape(r, 1),
                             w_update_ind=w_ind, h_update_ind=np.ones(r, bool), **nmf_kw)
    A = A[:, 0]

    # per-class reconstruction :158-202
    def class_ranges(num, ranks, lo, hi):
        out = []
        for i in range(num):
            a = lo + ranks[i] - 1
            b_ = hi if i == num - 1 else lo + ranks[i + 1] - 1
            out.append(slice(a, b_))
        return out
    ev = class_ranges(p['EVENT_NUM'], p['EVENT_RANK'], 0, R_x)
    no = class_ranges(p['NOISE_NUM'], p['NOISE_RANK'], R_x, R_x + R_d)
    if mel and p['MelConv']:
        Xm_hat = [g.melmat.T @ (B_Mel[:, s] @ A[s]) for s in ev]
        Dm_hat = [g.melmat.T @ (B_Mel[:, s] @ A[s]) for s in no]
        Ym_Mel_DFT = g.melmat.T @ Ym_Mel                         # :205-211
    else:
        Xm_hat = [B_DFT[:, s] @ A[s] for s in ev]
        Dm_hat = [B_DFT[:, s] @ A[s] for s in no]
        Ym_Mel_DFT = Ym
    Xm_hat_sum = np.sum(Xm_hat, axis=0)
    Dm_hat_sum = np.sum(Dm_hat, axis=0)

    # block sparsity :213-218
    if p['blk_sparse']:
        Q, g.r_blk = blk_sparse(Xm_hat_sum, Dm_hat_sum, g.r_blk, l, p)
    else:
        Q = np.ones(n2)

    # 3) gain :220-260
    lambda_dav = g.lambda_dav
    if l == 1:
        lambda_dav = Ym_Mel_DFT.copy()
    A_d_mag = np.sum(A[R_x:R_x + R_d]) / R_d
    A_x_mag = np.sum(A[:R_x]) / R_x
    beta = 20 * np.log10(A_d_mag / A_x_mag) * p['beta']
    if beta < p['beta']:
        beta = p['beta']
    elif beta >= p['beta_max']:
        beta = p['beta_max']
    lambda_dav = p['alpha_d'] * lambda_dav + (1 - p['alpha_d']) * Dm_hat_sum * beta
    lambda_d = lambda_dav
    if p['ENHANCE_METHOD'] == 'Wiener':
        G = Xm_hat_sum / (Xm_hat_sum + Dm_hat_sum)
    else:
        eta = (p['alpha_eta'] * g.Xm_tilde + (1 - p['alpha_eta']) * Xm_hat_sum * Q) / np.maximum(lambda_d, flr)
        eta = np.maximum(0.0031, eta)
        G = eta / (eta + 1.0)
    G = np.minimum(G, 1.0)
    if l <= p['init_N_len']:
        G = np.zeros(n2) + flr
        A_x_mag = flr
    Xm_tilde = G * Ym

    # 4) adaptation :262-347
    Q_control = (1 - np.mean(Q)) * p['Ar_up']
    gated = bool(p['adapt_train_N'] and (Q_control * A_d_mag > A_x_mag))
    w_iters = 0
    R_a_up = 0
    if gated:
        if l <= p['init_N_len']:
            D_ref = Ym
        else:
            M_ref = 1 - G
            M_ref[:p['DCbin']] = flr
            D_ref = Ym * M_ref
        g.lambda_d_blk = np.concatenate([g.lambda_d_blk[:, 1:p['m_a']], D_ref[:, None]], axis=1)
        g.Ad_blk = np.concatenate([g.Ad_blk[:, 1:p['m_a']], A[R_x:R_x + p['R_a'], None]], axis=1)
        r_up = Q_control * np.mean(g.Ad_blk, axis=1) > A_x_mag
        Ad_up = g.Ad_blk[r_up, :]
        if not np.all(np.any(Ad_up != 0, axis=1)):
            raise FloatingPointError('all-zero activation row in Ad_blk_up (reference dimension mismatch)')
        if g.update_switch == math.floor(p['overlap_m_a'] * p['m_a']):
            if mel:
                lam_blk = g.melmat @ g.lambda_d_blk
                Bsrc = B_Mel_d
            else:
                lam_blk = g.lambda_d_blk
                Bsrc = B_DFT_d
            B_up = Bsrc[:, :p['R_a']][:, r_up]
            B_rem = Bsrc[:, :p['R_a']][:, ~r_up]
            B_fix = B_Mel_d[:, p['R_a']:]                         # :307 / :328 [sic]
            R_a_up = B_up.shape[1]
            if R_a_up > 0:
                B_tmp, _, obj_w = sparse_nmf(lam_blk, init_w=B_up, init_h=Ad_up,
                                             w_update_ind=np.ones(R_a_up, bool),
                                             h_update_ind=np.zeros(R_a_up, bool), **nmf_kw)
                w_iters = obj_w['iters']
                newB = np.concatenate([B_rem, B_tmp, B_fix], axis=1)
            else:
                newB = np.concatenate([B_rem, B_fix], axis=1)
            if mel:
                g.B_Mel_d = newB
            else:
                g.B_DFT_d = newB
            g.update_switch = 1
        else:
            g.update_switch += 1

    # ISTFT :349-363
    ist = lambda mag: synth_ifft_buff(mag, Yp, p['framelength'], p['fftlength'], p['win_ISTFT'],
                                      p['preemph'], p['DCbin_back'], p['pow'])[:, 0] * p['overlapscale']
    x_tilde = ist(Xm_tilde)
    x_hat_i = d_hat_i = None
    if want_aux:
        x_hat_i = np.stack([ist(x) for x in Xm_hat])
        d_hat_i = np.stack([ist(d) for d in Dm_hat])

    g.Ym, g.Yp = Ym, Yp
    g.lambda_dav = lambda_dav
    g.Xm_tilde = Xm_tilde
    g.dbg = dict(A=A, h_iters=obj_h['iters'], h_cost=obj_h['cost'][-1] if len(obj_h['cost']) else 0.0,
                 Q=Q, G=G, gated=gated, R_a_up=R_a_up, w_iters=w_iters,
                 Xm_hat=Xm_hat_sum, Dm_hat=Dm_hat_sum, beta=beta)
    return x_hat_i, d_hat_i, x_tilde, g


# ----------------------------------------------------------------------------
# file-level drivers
# ----------------------------------------------------------------------------
def to_int16(x):
    """fwrite(...,'int16') of doubles: round half away from zero, saturate."""
    x = np.asarray(x, dtype=np.float64)
    r = np.sign(x) * np.floor(np.abs(x) + 0.5)
    return np.clip(r, -32768, 32767).astype(np.int16)


def num_hops(n_samples: int, p) -> int:
    """Hops the loop of filewise_run_IS16.m:102-169 executes: every full hop
    plus p.delay+1 flush iterations."""
    return n_samples // p['frameshift'] + p['delay'] + 1


def enhance_utterance(pcm, p, B_x, B_d, *, h_init, Ad_blk_init, A_d_init=None, B_Mel_x=None, B_Mel_d=None,
                      trace: Optional[list] = None, max_hops=None):
    """The hop loop of filewise_run_IS16.m:39-51,83-169 on an int16 signal
    (the samples after the 44-byte header).  Returns (int16 output, final g).
    In DFT mode the "Mel" slots hold the DFT bases (filewise_run_IS16.m:46-51)."""
    pcm = np.asarray(pcm)
    B_x = np.asarray(B_x, dtype=np.float64)
    B_d = np.asarray(B_d, dtype=np.float64)
    if B_d.shape[1] < p['R_d']:                                   # :39-43
        B_d = np.concatenate([B_d, B_d[:, :p['R_d'] - B_d.shape[1]]], axis=1)
        if B_Mel_d is not None:
            B_Mel_d = np.concatenate([B_Mel_d, B_Mel_d[:, :p['R_d'] - B_Mel_d.shape[1]]], axis=1)
    if p['B_sep_mode'] == 'Mel':
        B1_x, B1_d = B_Mel_x, B_Mel_d
    else:
        B1_x, B1_d = B_x, B_d
    g = init_buff(B1_x, B1_d, B_x, B_d, p, Ad_blk_init=Ad_blk_init, A_d_init=A_d_init)
    fl, fs_ = p['framelength'], p['frameshift']
    y = np.zeros(fl)
    ola = np.zeros(fl)
    out = []
    n_full = len(pcm) // fs_
    total = n_full + p['delay'] + 1
    if max_hops is not None:
        total = min(total, max_hops)
    for l in range(1, total + 1):
        if l <= n_full:
            y[:fl - fs_] = y[fs_:].copy()
            y[fl - fs_:] = pcm[(l - 1) * fs_:l * fs_]
        else:
            y = np.zeros(fl)                                      # :111-113
        _, _, d_frame, g = bnmf_sep_event_RT_IS16(y, l, g, p, h_init=h_init)
        if trace is not None:
            trace.append(dict(l=l, x_tilde=d_frame.copy(), Xm_tilde=g.Xm_tilde.copy(), Ym=g.Ym.copy(), **g.dbg))
        if l > p['delay']:                                        # :146,162-165
            ola[:fl - fs_] = ola[fs_:].copy()
            ola[fl - fs_:] = 0.0
            ola = ola + d_frame
            out.append(to_int16(ola[:fs_]))
    out = np.concatenate(out) if out else np.zeros(0, np.int16)
    return out, g


def enhance_chain(pcms, p, B_x, B_d, *, h_init, Ad_blk_inits, **kw):
    """A target-directory chain: src/NTF_sep_event_RT.m:28-38,136-139 hands
    the adapted noise basis of file i to file i+1 through B_D_u.mat (both the
    DFT and the "Mel" slot are overwritten by the load)."""
    outs = []
    Bd = np.asarray(B_d, dtype=np.float64)
    for pcm, ad in zip(pcms, Ad_blk_inits):
        out, g = enhance_utterance(pcm, p, B_x, Bd, h_init=h_init, Ad_blk_init=ad, **kw)
        outs.append(out)
        if p['adapt_train_N']:
            Bd = g.B_DFT_d
    return outs, Bd


# ----------------------------------------------------------------------------
# offline dictionary training core  (run_basis_train.m:60-63,80-91,112-116)
# ----------------------------------------------------------------------------
def basis_train_core(TF_pow, R, sample_idx, p, *, h_init):
    """sparse_nmf W+H on a power spectrogram with exemplar init
    (run_basis_train.m:80-88) then column normalisation + 1e-9 (:112-114).
    ``sample_idx`` replaces ``randsample`` (:81); ``h_init`` is rand(R,T)."""
    B_init = TF_pow[:, sample_idx]
    w, h, obj = sparse_nmf(TF_pow, init_w=B_init, init_h=h_init,
                           w_update_ind=np.ones(R, bool), h_update_ind=np.ones(R, bool),
                           max_iter=p['max_iter'], sparsity=p['sparsity'], conv_eps=p['conv_eps'],
                           cf=p['cf'], cost_check=bool(p['cost_check']))
    wn = np.sqrt(np.sum(w ** 2, axis=0))
    return w / wn + 1e-9, h, obj


def training_spectrogram(s_full, p, DCbin=None):
    """run_basis_train.m:60-63: stft_fft magnitude -> drop all-zero columns
    -> (Splice=0) -> .^pow + nonzerofloor."""
    mag, _ = stft_fft(s_full, p['framelength'], p['frameshift'], p['fftlength'],
                      p['DCbin'] if DCbin is None else DCbin, p['win_STFT'], p['preemph'])
    mag = mag[:, np.any(mag != 0, axis=0)]
    return mag ** p['pow'] + p['nonzerofloor']


# ----------------------------------------------------------------------------
# I/O helpers and RNG stand-ins
# ----------------------------------------------------------------------------
def read_wav_pcm(path):
    """filewise_run_IS16.m:92-97: skip 22 int16 (44-byte header), rest is PCM."""
    raw = np.fromfile(path, dtype='<i2')
    return raw[22:]


def park_miller(n, state=1):
    """Park-Miller minimal standard LCG (a=16807, m=2^31-1): documented
    candidate for MATLAB's legacy rand('seed',.) stream."""
    out = np.empty(n)
    s = state
    for i in range(n):
        s = (16807 * s) % 2147483647
        out[i] = s / 2147483647.0
    return out


def default_rng_inputs(p, R=None, with_A_d=False):
    """Deterministic stand-ins for the MATLAB RNG draws (see tests/golden).

    init_buff.m:37-38 draws ``g.A_d = rand(R_d, m)`` BEFORE ``g.Ad_blk = rand(R_a, m_a)``; the stream keeps that order
    (the A_d values are overwritten before their first use, so only their position in the stream matters).
    tests/golden/ref_shadow/rand.m replays the same stream / the same Park-Miller sequence inside Octave or MATLAB."""
    R = p['R_x'] + p['R_d'] if R is None else R
    h_init = park_miller(R, 1)
    rs = np.random.RandomState(5489)
    A_d = rs.rand(p['R_d'])
    Ad = rs.rand(p['m_a'], p['R_a']).T.copy()
    if with_A_d:
        return h_init, Ad, A_d
    return h_init, Ad