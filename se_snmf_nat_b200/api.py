"""Host-side mirror of the reference's MATLAB interface for the SNMF-NAT hot path.

Every function here has the name, argument meaning and error behaviour of the
reference function it replaces (cited as file:line of lordet01/SE_SNMF_NAT) and
is a thin marshalling layer over the C ABI of ``libsnmfnat.so``
(``include/snmfnat.h``) -- the same calls the MEX gateways in ``mex/`` make.
Nothing is computed on the CPU here; without the CUDA library the calls raise.

``p`` is a plain dict with the field names of the reference's ``global p``
(settings/initial_setting_SNMF_NAT.m); see ``settings.load_settings``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import BatchStats, NmfOpts, Params, SnmfnatError, check

_CF = {"is": _lib.CF_IS, "kl": _lib.CF_KL, "ed": _lib.CF_ED}


def _dptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, order="F") -> np.ndarray:
    """MATLAB matrices cross the boundary as column-major doubles."""
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])


def sqrt_hann_periodic(n: int) -> np.ndarray:
    """sqrt(hann(n,'periodic')), settings/initial_setting_SNMF_NAT.m:33,35."""
    k = np.arange(n, dtype=np.float64)
    return np.sqrt(0.5 * (1.0 - np.cos(2.0 * np.pi * k / n)))


def default_p() -> dict:
    """The shipped parameter set (settings/initial_setting_SNMF_NAT.m) as the library reports it."""
    lib = _lib.load()
    ps = Params()
    lib.snmfnat_params_default(C.byref(ps))
    p = {}
    for name, ctype in Params._fields_:
        if name.startswith("reserved"):
            continue
        v = getattr(ps, name)
        p[name] = list(v) if hasattr(v, "__len__") else v
    p["EVENT_RANK"] = p["EVENT_RANK"][:p["EVENT_NUM"]]
    p["NOISE_RANK"] = p["NOISE_RANK"][:p["NOISE_NUM"]]
    p["cf"] = "kl"
    p["ENHANCE_METHOD"] = "MMSE"
    p["B_sep_mode"] = "DFT"
    p["win_STFT"] = sqrt_hann_periodic(p["framelength"])
    p["win_ISTFT"] = sqrt_hann_periodic(p["framelength"])
    p["random_seed"] = 1
    p["display"] = 0
    p["useGPU"] = 0
    return p


def params_struct(p: dict) -> Params:
    """Flatten the fields of ``p`` the hot path reads into struct snmfnat_params.
    Unknown fields are ignored, missing ones keep the shipped defaults."""
    lib = _lib.load()
    ps = Params()
    lib.snmfnat_params_default(C.byref(ps))
    for name, ctype in Params._fields_:
        if name.startswith("reserved") or name not in p:
            continue
        v = p[name]
        if name in ("EVENT_RANK", "NOISE_RANK"):
            vals = list(np.atleast_1d(v).astype(int))
            if len(vals) > _lib.MAX_CLASSES:
                raise ValueError(f"{name} has more than {_lib.MAX_CLASSES} classes")
            arr = getattr(ps, name)
            for i in range(_lib.MAX_CLASSES):
                arr[i] = vals[i] if i < len(vals) else 0
        elif name == "cf":
            if isinstance(v, str):
                if v in _CF:
                    ps.cf = _CF[v]
                else:  # sparse_nmf.m:106-109: any other string keeps p.beta
                    ps.cf = _lib.CF_BETA
                    ps.beta_div = float(p.get("beta_div", 1.0))
            else:
                ps.cf = int(v)
        elif name == "ENHANCE_METHOD":
            ps.ENHANCE_METHOD = (_lib.ENH_WIENER if v == "Wiener" else _lib.ENH_MMSE) if isinstance(v, str) else int(v)
        elif name == "B_sep_mode":
            ps.B_sep_mode = (_lib.SEP_MEL if v == "Mel" else _lib.SEP_DFT) if isinstance(v, str) else int(v)
        elif ctype in (C.c_int32,):
            setattr(ps, name, int(v))
        else:
            setattr(ps, name, float(np.asarray(v).reshape(-1)[0]))
    return ps


class Context:
    """One CUDA device (snmfnat_ctx)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.snmfnat_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.snmfnat_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self._lib.snmfnat_ctx_sync(self._h))

    @property
    def cuda_stream(self) -> int:
        return int(self._lib.snmfnat_ctx_cuda_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.snmfnat_ctx_launch_count(self._h))


_default_ctx = {}


def get_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Batch:
    """A batch of utterances enhanced in lock step on one GPU (snmfnat_batch):
    the hop loop of filewise_run_IS16.m:86-169 for every utterance."""

    def __init__(self, ctx: Context, p: dict, B_x, B_d, lengths: Sequence[int], h_init, Ad_blk_init,
                 chain_id: Optional[Sequence[int]] = None):
        self.ctx = ctx
        self._lib = ctx._lib
        self.p = p
        ps = params_struct(p)
        B_x = _f64(B_x)
        B_d = _f64(B_d)
        if B_d.shape[1] < ps.R_d:  # filewise_run_IS16.m:39-43
            B_d = _f64(np.concatenate([B_d, B_d[:, :ps.R_d - B_d.shape[1]]], axis=1))
        if B_x.shape[1] != ps.R_x or B_d.shape[1] != ps.R_d or B_x.shape[0] != B_d.shape[0]:
            raise ValueError("basis shapes do not match p.R_x / p.R_d")
        self.n_utt = len(lengths)
        self.lengths = np.asarray(lengths, dtype=np.int64)
        win_s = _f64(np.asarray(p.get("win_STFT", sqrt_hann_periodic(ps.framelength))).ravel())
        win_i = _f64(np.asarray(p.get("win_ISTFT", sqrt_hann_periodic(ps.framelength))).ravel())
        if win_s.size != ps.framelength or win_i.size != ps.framelength:
            raise ValueError("window length must equal p.framelength")
        h_init = _f64(np.asarray(h_init).ravel())
        if h_init.size != ps.R_x + ps.R_d:
            raise ValueError("h_init must have R_x+R_d entries")
        ad = None
        stride = 0
        if Ad_blk_init is not None:
            ad = np.asarray(Ad_blk_init, dtype=np.float64)
            if ad.ndim == 2:
                if ad.shape != (ps.R_a, ps.m_a):
                    raise ValueError("Ad_blk_init must be R_a x m_a")
                ad = np.asfortranarray(ad)
            else:
                if ad.shape != (self.n_utt, ps.R_a, ps.m_a):
                    raise ValueError("Ad_blk_init must be n_utt x R_a x m_a")
                # per utterance column-major R_a x m_a blocks, back to back
                ad = np.ascontiguousarray(np.transpose(ad, (0, 2, 1)))
                stride = ps.R_a * ps.m_a
        chain = None
        if chain_id is not None:
            chain = np.asarray(chain_id, dtype=np.int32)
        h = C.c_void_p()
        check(self._lib.snmfnat_batch_create(
            ctx._h, C.byref(ps), _dptr(win_s), _dptr(win_i), _dptr(B_x), _dptr(B_d), B_x.shape[0], self.n_utt,
            self.lengths.ctypes.data_as(C.POINTER(C.c_int64)),
            chain.ctypes.data_as(C.POINTER(C.c_int32)) if chain is not None else None,
            _dptr(h_init), _dptr(ad), stride, C.byref(h)))
        self._h = h
        self._ps = ps
        self.out_lengths = np.array([self._lib.snmfnat_batch_out_len(h, u) for u in range(self.n_utt)], dtype=np.int64)
        self.total_hops = int(self._lib.snmfnat_batch_total_hops(h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.snmfnat_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, pcms: Sequence[np.ndarray]):
        arrs = [np.ascontiguousarray(x, dtype=np.int16) for x in pcms]
        for a, n in zip(arrs, self.lengths):
            if a.size != n:
                raise ValueError("PCM length differs from the length given at creation")
        ptrs = (C.POINTER(C.c_int16) * self.n_utt)(*[a.ctypes.data_as(C.POINTER(C.c_int16)) for a in arrs])
        check(self._lib.snmfnat_batch_upload(self._h, ptrs))

    def upload_packed(self, packed_ptr: int):
        """``packed_ptr``: address of a (pinned) host buffer with the utterances back to back."""
        check(self._lib.snmfnat_batch_upload_packed(self._h, C.cast(packed_ptr, C.POINTER(C.c_int16))))

    def enable_trace(self, on=True):
        check(self._lib.snmfnat_batch_enable_trace(self._h, 1 if on else 0))

    def run(self):
        check(self._lib.snmfnat_batch_run(self._h))

    def download(self):
        outs = [np.empty(int(n), dtype=np.int16) for n in self.out_lengths]
        ptrs = (C.POINTER(C.c_int16) * self.n_utt)(*[a.ctypes.data_as(C.POINTER(C.c_int16)) for a in outs])
        check(self._lib.snmfnat_batch_download(self._h, ptrs))
        return outs

    def download_packed(self, packed_ptr: int):
        check(self._lib.snmfnat_batch_download_packed(self._h, C.cast(packed_ptr, C.POINTER(C.c_int16))))

    def set_profile(self, on=True):
        check(self._lib.snmfnat_batch_set_profile(self._h, 1 if on else 0))

    def profile(self) -> dict:
        """Device time (ms) per kernel class of the last run, from CUDA events on the launch stream."""
        ms = (C.c_double * 6)()
        cnt = (C.c_int64 * 6)()
        check(self._lib.snmfnat_batch_get_profile(self._h, ms, cnt))
        names = ["stft", "hsolve", "gain", "wsolve", "istft", "total"]
        return {n: {"ms": ms[i], "launches": int(cnt[i])} for i, n in enumerate(names)}

    def stats(self) -> dict:
        s = BatchStats()
        check(self._lib.snmfnat_batch_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in BatchStats._fields_ if k != "reserved"}

    def trace(self, u: int, what: str) -> np.ndarray:
        nh = int(self.lengths[u] // self._ps.frameshift + self._ps.delay + 1)
        F = self._ps.fftlength // 2 + 1
        width = {"A": self._ps.R_x + self._ps.R_d, "Q": F, "G": F, "Xm_tilde": F, "Ym": F}.get(what, 1)
        buf = np.empty((nh, width), dtype=np.float64)
        check(self._lib.snmfnat_batch_get_trace(self._h, int(u), what.encode(), _dptr(buf), buf.size))
        return buf if width > 1 else buf[:, 0]

    def noise_basis(self, u: int) -> np.ndarray:
        F = self._ps.fftlength // 2 + 1
        out = np.empty((F, self._ps.R_d), dtype=np.float64, order="F")
        check(self._lib.snmfnat_batch_get_noise_basis(self._h, int(u), _dptr(out)))
        return out


def enhance_batch(pcms: Sequence[np.ndarray], p: dict, B_x, B_d, *, h_init, Ad_blk_init, device: int = 0,
                  chain_id=None, return_stats=False):
    """Enhance a list of int16 signals (samples after the 44-byte WAV header) exactly like running
    filewise_run_IS16.m on each of them; returns the int16 outputs."""
    ctx = get_context(device)
    b = Batch(ctx, p, B_x, B_d, [len(x) for x in pcms], h_init, Ad_blk_init, chain_id)
    try:
        b.upload(pcms)
        b.run()
        outs = b.download()
        if return_stats:
            return outs, b.stats()
        return outs
    finally:
        b.close()


def filewise_run_IS16(path_in: str, path_denoise: str, p: dict, B_DFT_x, B_DFT_d, *, h_init, Ad_blk_init,
                      device: int = 0):
    """filewise_run_IS16.m:54-186 for one file: read the int16 samples after the 44-byte header (:92-97),
    enhance, write raw PCM wrapped as a 16-bit WAV (src/pcm2wav.m)."""
    raw = np.fromfile(path_in, dtype="<i2")
    pcm = raw[22:]
    out = enhance_batch([pcm], p, B_DFT_x, B_DFT_d, h_init=h_init, Ad_blk_init=Ad_blk_init, device=device)[0]
    import wave
    with wave.open(path_denoise, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(p.get("fs", 16000)))
        w.writeframes(out.astype("<i2").tobytes())
    return out
