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
                 chain_id: Optional[Sequence[int]] = None, B_Mel_x=None, B_Mel_d=None, melmat=None):
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
        self.n1 = 0
        if ps.B_sep_mode == _lib.SEP_MEL:   # filewise_run_IS16.m:46-51: the Mel slots take B_Mel_sub
            if B_Mel_x is None or B_Mel_d is None:
                raise ValueError("B_sep_mode='Mel' needs B_Mel_x and B_Mel_d")
            bmx, bmd = _f64(B_Mel_x), _f64(B_Mel_d)
            if bmd.shape[1] < ps.R_d:
                bmd = _f64(np.concatenate([bmd, bmd[:, :ps.R_d - bmd.shape[1]]], axis=1))
            if bmx.shape != (bmd.shape[0], ps.R_x) or bmd.shape[1] != ps.R_d:
                raise ValueError("Mel basis shapes do not match p.R_x / p.R_d")
            mm = None
            if melmat is not None:
                mm = _f64(melmat)
                if mm.shape != (B_x.shape[0], bmx.shape[0]):
                    raise ValueError("melmat must be n2 x n1 (what mel_matrix.m returns)")
            check(self._lib.snmfnat_batch_set_mel(h, _dptr(bmx), _dptr(bmd), bmx.shape[0], _dptr(mm)))
            self.n1 = bmx.shape[0]
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
        """``packed_ptr``: address of a (pinned) host buffer with the utterances back to back.  Asynchronous: the buffer
        must stay valid and unmodified until the next synchronising call (``ctx.sync()``, ``download*``)."""
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

    def set_groups(self, n: int):
        """Scheduling only: n interleaved slot groups on separate CUDA streams (snmfnat_batch_set_groups)."""
        check(self._lib.snmfnat_batch_set_groups(self._h, int(n)))

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
        F = self.n1 if self.n1 else self._ps.fftlength // 2 + 1   # Mel mode adapts B_Mel_d
        out = np.empty((F, self._ps.R_d), dtype=np.float64, order="F")
        check(self._lib.snmfnat_batch_get_noise_basis(self._h, int(u), _dptr(out)))
        return out


def mel_matrix(fs: int, NbCh: int, Nfft: int, warp: float = 1.0, fhigh: Optional[float] = None) -> np.ndarray:
    """M = mel_matrix(fs,NbCh,Nfft,warp,fhigh), src/mel_matrix.m:9-38 (dense (Nfft/2+1) x NbCh)."""
    M = np.zeros((Nfft // 2 + 1, NbCh), dtype=np.float64, order="F")
    check(_lib.load().snmfnat_mel_matrix(int(fs), int(NbCh), int(Nfft), float(warp),
                                         float(fhigh if fhigh is not None else -1.0), _dptr(M)))
    return M


def enhance_batch(pcms: Sequence[np.ndarray], p: dict, B_x, B_d, *, h_init, Ad_blk_init, device: int = 0,
                  chain_id=None, return_stats=False, B_Mel_x=None, B_Mel_d=None, melmat=None):
    """Enhance a list of int16 signals (samples after the 44-byte WAV header) exactly like running
    filewise_run_IS16.m on each of them; returns the int16 outputs."""
    ctx = get_context(device)
    b = Batch(ctx, p, B_x, B_d, [len(x) for x in pcms], h_init, Ad_blk_init, chain_id, B_Mel_x=B_Mel_x, B_Mel_d=B_Mel_d,
              melmat=melmat)
    try:
        b.upload(pcms)
        b.run()
        outs = b.download()
        if return_stats:
            return outs, b.stats()
        return outs
    finally:
        b.close()


def enhance_batch_multi(pcms: Sequence[np.ndarray], p: dict, B_x, B_d, *, h_init, Ad_blk_init, devices: Sequence[int],
                        chain_id=None):
    """The same corpus on several GPUs of one node from this one process (snmfnat_enhance_batch_multi): utterances /
    chains are split over ``devices`` longest first, one host thread per device, no collective."""
    lib = _lib.load()
    ps = params_struct(p)
    arrs = [np.ascontiguousarray(x, dtype=np.int16) for x in pcms]
    n = len(arrs)
    lens = np.array([a.size for a in arrs], dtype=np.int64)
    hop, delay = int(p["frameshift"]), int(p["delay"])
    outs = [np.empty(int((ln // hop + delay + 1 - delay) * hop), dtype=np.int16) for ln in lens]
    i16p = C.POINTER(C.c_int16)
    pin = (i16p * n)(*[a.ctypes.data_as(i16p) for a in arrs])
    pout = (i16p * n)(*[a.ctypes.data_as(i16p) for a in outs])
    Bx, Bd = _f64(B_x), _f64(B_d)
    ad = np.ascontiguousarray(np.stack([_f64(a).ravel(order="F") for a in np.broadcast_to(
        np.asarray(Ad_blk_init, dtype=np.float64), (n,) + np.asarray(Ad_blk_init).shape[-2:])]))
    ch = None if chain_id is None else np.ascontiguousarray(chain_id, dtype=np.int32)
    dev = (C.c_int * len(devices))(*[int(d) for d in devices])
    ws, wi, h0 = _f64(p["win_STFT"]).ravel(), _f64(p["win_ISTFT"]).ravel(), _f64(h_init).ravel()
    check(lib.snmfnat_enhance_batch_multi(dev, len(devices), C.byref(ps), _dptr(ws), _dptr(wi), _dptr(Bx), _dptr(Bd),
                                          Bx.shape[0], n, pin, lens.ctypes.data_as(C.POINTER(C.c_int64)),
                                          None if ch is None else ch.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(h0),
                                          _dptr(ad), ad.shape[1], pout))
    return outs


def save_basis_mat(path: str, B_DFT_sub, B_Mel_sub, A_DFT_sub=None, A_Mel_sub=None):
    """run_basis_train.m:136 saves 'B_DFT_sub', 'B_Mel_sub', 'A_DFT_sub', 'A_Mel_sub' with '-v7.3' (HDF5).  This writes the
    same four variables as MAT v5, which the reference's own loader (`load(...)`, run_basis_train.m:138, filewise_run_IS16.m:24-37)
    reads just the same: `load` detects the version.  (HDF5 is not available in this environment; v5 caps a variable at
    2 GB, enough for the dictionaries and for the activations of ~2.6 M frames at R = 100.)"""
    import scipy.io
    d = {"B_DFT_sub": _f64(B_DFT_sub), "B_Mel_sub": _f64(B_Mel_sub)}
    d["A_DFT_sub"] = np.zeros((1, 1)) if A_DFT_sub is None else _f64(A_DFT_sub)      # run_basis_train.m:96-97: 0 when unused
    d["A_Mel_sub"] = np.zeros((1, 1)) if A_Mel_sub is None else _f64(A_Mel_sub)
    scipy.io.savemat(path, d, format="5", do_compression=False)


def load_basis_mat(path: str) -> dict:
    """The variables of a basis/*/R_<n>.mat file (MAT v5 / v7 as shipped by the reference)."""
    import scipy.io
    m = scipy.io.loadmat(path)
    return {k: np.asarray(v, dtype=np.float64) for k, v in m.items() if not k.startswith("__")}


def pcm2wav_samples(pcm: np.ndarray) -> np.ndarray:
    """src/pcm2wav.m:9-10: the raw int16 output is divided by 32767 and written with wavwrite(..., 16, ...), which
    quantises as round(x * 32768) (half away from zero) clipped to [-32768, 32767]: samples above 16383 in magnitude
    move by one LSB."""
    x = np.asarray(pcm, dtype=np.float64) / 32767.0 * 32768.0
    y = np.sign(x) * np.floor(np.abs(x) + 0.5)
    return np.clip(y, -32768, 32767).astype(np.int16)


def filewise_run_IS16(path_in: str, path_denoise: str, p: dict, B_DFT_x, B_DFT_d, *, h_init, Ad_blk_init,
                      device: int = 0):
    """filewise_run_IS16.m:54-186 for one file: read the int16 samples after the 44-byte header (:92-97), enhance,
    then src/pcm2wav.m: the PCM goes to a 16-bit WAV through wavwrite's x/32767*32768 re-quantisation.  Returns the raw
    int16 output of the frame loop (what fwrite puts into the file before pcm2wav, :165)."""
    raw = np.fromfile(path_in, dtype="<i2")
    pcm = raw[22:]
    out = enhance_batch([pcm], p, B_DFT_x, B_DFT_d, h_init=h_init, Ad_blk_init=Ad_blk_init, device=device)[0]
    import wave
    with wave.open(path_denoise, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(p.get("fs", 16000)))
        w.writeframes(pcm2wav_samples(out).astype("<i2").tobytes())
    return out


# ----------------------------------------------------------------------------------------------------------------
# L1 numeric functions (same names / argument meaning as the reference's src/*.m)
# ----------------------------------------------------------------------------------------------------------------
def _nmf_opts(p: dict, n: int, r: int, *, sparsity_key="sparsity", conv_key="conv_eps"):
    """Optional fields of p with the defaults of sparse_nmf.m:79-97; `cost_check` has no default there (:260)."""
    if "cost_check" not in p:
        raise KeyError("p.cost_check is required (sparse_nmf.m:260 reads it without a default)")
    o = NmfOpts()
    o.max_iter = int(p.get("max_iter", 100))
    cf = p.get("cf", "kl")
    if isinstance(cf, str):
        if cf in _CF:
            o.cf = _CF[cf]
        else:
            o.cf = _lib.CF_BETA
            o.beta_div = float(p.get("beta", 1.0))
    else:
        o.cf = int(cf)
    o.cost_check = int(bool(p["cost_check"]))
    o.conv_eps = float(p.get(conv_key, 0.0))
    o.precision = 0
    sp = np.asarray(p.get(sparsity_key, 0.0), dtype=np.float64)
    if sp.size == 1:
        sp = sp.reshape(1, 1)
    elif sp.ndim == 1 or sp.shape[1] == 1:
        sp = sp.reshape(r, 1)
    elif sp.shape != (r, n):
        raise ValueError("sparsity must be scalar, r x 1 or r x n")
    o.sparsity_rows, o.sparsity_cols = sp.shape
    return o, _f64(sp)


def _nmf_inits(v, p, rand):
    m, n = v.shape
    if "init_w" in p:
        w0 = np.array(p["init_w"], dtype=np.float64)
        r = w0.shape[1]
        if "r" in p and r < int(p["r"]):                       # sparse_nmf.m:125-127
            if rand is None:
                raise ValueError("p.r > size(init_w,2): pass rand= to supply MATLAB's rand(m, r-ri)")
            w0 = np.concatenate([w0, np.asarray(rand(m, int(p["r"]) - r), dtype=np.float64)], axis=1)
            r = int(p["r"])
    else:
        if "r" not in p:
            raise ValueError("Number of components or initialization must be given")   # sparse_nmf.m:118
        if rand is None:
            raise ValueError("no p.init_w: pass rand= to supply MATLAB's rand(m, r) (RNG stays on the host)")
        r = int(p["r"])
        w0 = np.asarray(rand(m, r), dtype=np.float64)
    ih = p.get("init_h", None)
    if ih is None:
        if rand is None:
            raise ValueError("no p.init_h: pass rand= to supply MATLAB's rand(r, n) (RNG stays on the host)")
        h0 = np.asarray(rand(r, n), dtype=np.float64)
    elif isinstance(ih, str) and ih == "ones":
        h0 = np.ones((r, n))
    else:
        h0 = np.array(ih, dtype=np.float64).reshape(r, n)
    return _f64(w0), _f64(h0), r


def _ind(p, key, r):
    if key not in p or p[key] is None:
        return None
    a = np.ascontiguousarray(np.asarray(p[key]).ravel().astype(np.uint8))
    if a.size != r:
        raise ValueError(f"{key} must have r entries")
    return a


def _u8ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def sparse_nmf(v, p: dict, *, rand=None, device: int = 0):
    """[w, h, objective] = sparse_nmf(v, p)     src/sparse_nmf.m:1-292.
    `rand(m, n)` stands in for MATLAB's rand after rand('seed', p.random_seed) when p.init_w / p.init_h are absent."""
    ctx = get_context(device)
    v = _f64(np.atleast_2d(np.asarray(v, dtype=np.float64)) if np.ndim(v) > 1 else np.asarray(v, dtype=np.float64).reshape(-1, 1))
    m, n = v.shape
    w0, h0, r = _nmf_inits(v, p, rand)
    o, sp = _nmf_opts(p, n, r)
    wi, hi = _ind(p, "w_update_ind", r), _ind(p, "h_update_ind", r)
    w = np.empty((m, r), order="F")
    h = np.empty((r, n), order="F")
    div = np.zeros(max(o.max_iter, 1))
    cost = np.zeros(max(o.max_iter, 1))
    its = C.c_int(0)
    check(ctx._lib.snmfnat_sparse_nmf(ctx._h, _dptr(v), m, n, r, C.byref(o), _dptr(sp), _dptr(w0), _dptr(h0), _u8ptr(wi),
                                      _u8ptr(hi), _dptr(w), _dptr(h), _dptr(div), _dptr(cost), C.byref(its)))
    k = its.value
    return w, h, {"div": div[:k], "cost": cost[:k], "iters": k}


def snmf_mdi(v, Dm, p: dict, *, rand=None, soft=False, device: int = 0):
    """[v_MDI, h, objective] = snmf_mdi(v, Dm, p)   src/snmf_mdi.m (binary mask) / src/snmf_mdi_Sm.m (soft=True).
    Reads p.sparsity_mdi / p.conv_eps_mdi (both effectively required, snmf_mdi.m:93-99)."""
    for k in ("sparsity_mdi", "conv_eps_mdi"):
        if k not in p:
            raise KeyError(f"p.{k} is required (the defaults at snmf_mdi.m:93-99 test the wrong field names)")
    ctx = get_context(device)
    v = _f64(np.asarray(v, dtype=np.float64).reshape(np.shape(v)[0], -1))
    m, n = v.shape
    mk = _f64(np.asarray(Dm, dtype=np.float64).reshape(m, n))
    w0, h0, r = _nmf_inits(v, p, rand)
    o, sp = _nmf_opts(p, n, r, sparsity_key="sparsity_mdi", conv_key="conv_eps_mdi")
    wi, hi = _ind(p, "w_update_ind", r), _ind(p, "h_update_ind", r)
    out = np.empty((m, n), order="F")
    h = np.empty((r, n), order="F")
    div = np.zeros(max(o.max_iter, 1))
    cost = np.zeros(max(o.max_iter, 1))
    its = C.c_int(0)
    check(ctx._lib.snmfnat_snmf_mdi(ctx._h, _dptr(v), _dptr(mk), 1 if soft else 0, m, n, r, C.byref(o), _dptr(sp),
                                    _dptr(w0), _dptr(h0), _u8ptr(wi), _u8ptr(hi), _dptr(out), _dptr(h), _dptr(div),
                                    _dptr(cost), C.byref(its)))
    k = its.value
    return out, h, {"div": div[:k], "cost": cost[:k], "iters": k}


def snmf_mdi_Sm(v, Sm, p: dict, **kw):
    """src/snmf_mdi_Sm.m: the soft-mask variant."""
    return snmf_mdi(v, Sm, p, soft=True, **kw)


def DNMF_adapt(Y, D, B, p: dict, *, rand, device: int = 0):
    """B_a = DNMF_adapt(Y, D, B, p)      src/DNMF_adapt.m:1-21 (needs p.R_x, p.R_d)."""
    ctx = get_context(device)
    Y = _f64(Y); D = _f64(D); B = _f64(B)
    F, n = Y.shape
    R_x, R_d = int(p["R_x"]), int(p["R_d"])
    o, sp = _nmf_opts(p, n, R_x + R_d)
    h0 = _f64(np.asarray(rand(R_x + R_d, n), dtype=np.float64))
    out = np.empty((F, R_d), order="F")
    check(ctx._lib.snmfnat_dnmf_adapt(ctx._h, _dptr(Y), _dptr(D), _dptr(B), F, n, R_x, R_d, C.byref(o), _dptr(sp),
                                      _dptr(h0), _dptr(out)))
    return out


def GIST_NTF(p: dict, B, S_mag, *, rand, A=None, variant_c: bool = False, device: int = 0):
    """[C, A] = GIST_NTF(p, B, S_mag)      src/GIST_NTF.m:1-160  (variant_c: src/GIST_NTF_C.m, objective only when
    p.cost_check).  S_mag is Channel x N x M, B N x K; `rand(Channel, K)` supplies the rand of :14.  Returns
    (C, A, objective) with A = ones(M, K) unless given."""
    ctx = get_context(device)
    S = np.asfortranarray(np.asarray(S_mag, dtype=np.float64))
    Ch, N, M = S.shape
    B = _f64(B)
    K = B.shape[1]
    if B.shape[0] != N:
        raise ValueError("B must be N x K with N = size(S_mag, 2)")
    C0 = _f64(np.asarray(rand(Ch, K), dtype=np.float64))
    Aa = _f64(A) if A is not None else None
    max_iter = int(p.get("max_iter", 100))
    Cout = np.zeros((Ch, K), order="F")
    div = np.zeros(max(max_iter, 1))
    cost = np.zeros(max(max_iter, 1))
    its = C.c_int(0)
    cc = int(bool(p.get("cost_check", 1))) if variant_c else -1
    check(ctx._lib.snmfnat_gist_ntf(ctx._h, _dptr(S), Ch, N, M, _dptr(B), K, _dptr(C0), _dptr(Aa), float(p.get("sparsity", 0.0)),
                                    float(p.get("nonzerofloor", 1e-9)), max_iter, float(p.get("conv_eps", 0.0)), cc,
                                    _dptr(Cout), _dptr(div), _dptr(cost), C.byref(its)))
    k = its.value if cc != 0 else 0
    return Cout, (Aa if Aa is not None else np.ones((M, K))), {"div": div[:k], "cost": cost[:k], "iters": its.value}


def tf_features(S_mag, pow_: float, floor_: float, melmat=None, *, device: int = 0):
    """S_mag.^pow + floor, optionally projected on a filterbank (melmat' * X) -- the feature step of
    run_basis_train.m:63,70-78 / run_basis_DNMF(_Mel).m."""
    ctx = get_context(device)
    S = _f64(S_mag)
    F, T = S.shape
    mm = _f64(melmat) if melmat is not None else None
    n1 = mm.shape[1] if mm is not None else 0
    if mm is not None and mm.shape[0] != F:
        raise ValueError("melmat must be F x n1")
    out = np.zeros((n1 if mm is not None else F, T), order="F")
    check(ctx._lib.snmfnat_tf_features(ctx._h, _dptr(S), F, T, float(pow_), float(floor_), _dptr(mm), n1, _dptr(out)))
    return out


def _dnmf_features(sig, p, melmat, device):
    S, _ = stft_fft(sig, p["framelength"], p["frameshift"], p["fftlength"], p["DCbin"], p["win_STFT"], p["preemph"],
                    device=device)
    S = S[:, np.any(S != 0, axis=0)]                      # X = X(:,any(X,1))
    if int(p.get("Splice", 0)) != 0:
        raise SnmfnatError(-4, "frame_splice with Splice > 0 is not implemented")
    return tf_features(S, p["pow"], p["nonzerofloor"], melmat, device=device)


def run_basis_DNMF(x, d, B, p: dict, *, rand=None, mel: bool = False, device: int = 0):
    """B_hat = run_basis_DNMF(x, d, B, p)         run_basis_DNMF.m:1-57  (mel=False)
       B_hat = run_basis_DNMF_Mel(x, d, B, p)     run_basis_DNMF_Mel.m:1-95 (mel=True)
    Discriminative retraining of [B_x B_d] on a clean / noise pair: activations of the mixture with the dictionary
    fixed (Eq. 6), then W-only updates of the speech atoms on the clean spectrogram and of the noise atoms on the
    noise spectrogram with those activations (Eq. 7).  Every numeric step runs on the GPU (STFT, features, the three
    sparse_nmf solves).  `rand(m, n)` supplies rand(r, T) of sparse_nmf.m:134 for the first solve."""
    x = np.asarray(x, dtype=np.float64).ravel()
    d = np.asarray(d, dtype=np.float64).ravel()
    n = min(x.size, d.size)
    x, d = x[:n], d[:n]
    y = x + d
    melmat = mel_matrix(p["fs"], p["F_order"], p["fftlength"], 1.0, p["fs"] / 2) if mel else None
    X = _dnmf_features(x, p, melmat, device)
    D = _dnmf_features(d, p, melmat, device)
    Y = _dnmf_features(y, p, melmat, device)
    R_x, R_d = int(p["R_x"]), int(p["R_d"])
    B = _f64(B)
    q = dict(p)
    q.update(w_update_ind=np.zeros(R_x + R_d, bool), h_update_ind=np.ones(R_x + R_d, bool), init_w=B)
    q.pop("init_h", None)
    _, A_hat, _ = sparse_nmf(Y, q, rand=rand, device=device)
    q.update(w_update_ind=np.ones(R_x, bool), h_update_ind=np.zeros(R_x, bool), init_w=B[:, :R_x], init_h=A_hat[:R_x, :],
             r=R_x)
    B_hat_x, _, _ = sparse_nmf(X, q, device=device)
    q.update(w_update_ind=np.ones(R_d, bool), h_update_ind=np.zeros(R_d, bool), init_w=B[:, R_x:R_x + R_d],
             init_h=A_hat[R_x:R_x + R_d, :], r=R_d)
    B_hat_d, _, _ = sparse_nmf(D, q, device=device)
    return np.concatenate([B_hat_x, B_hat_d], axis=1)


def run_basis_DNMF_Mel(x, d, B, p: dict, **kw):
    return run_basis_DNMF(x, d, B, p, mel=True, **kw)


def stft_fft(s, sz, shift, fftlen, DCbin, win, preemph, *, device: int = 0):
    """[S_mag, S_phase] = stft_fft(s, sz, shift, fftlen, DCbin, win, preemph)     src/stft_fft.m:1-37."""
    ctx = get_context(device)
    s = _f64(np.asarray(s, dtype=np.float64).ravel())
    win = _f64(np.asarray(win, dtype=np.float64).ravel())
    if win.size != sz:
        raise ValueError("window length must equal sz")
    half = fftlen // 2 + 1
    nfr = s.size // shift
    mag = np.zeros((half, nfr), order="F")
    ph = np.zeros((half, nfr), order="F")
    check(ctx._lib.snmfnat_stft_fft(ctx._h, _dptr(s), s.size, int(sz), int(shift), int(fftlen), int(DCbin), _dptr(win),
                                    float(preemph), _dptr(mag), _dptr(ph)))
    return mag, ph


def synth_ifft_buff(TF_mag, TF_phase, sz, fftlen, win, preemph, DCbin_back, pow_, *, device: int = 0):
    """s_buff = synth_ifft_buff(TF_mag, TF_phase, sz, fftlen, win, preemph, DCbin_back, pow)  src/synth_ifft_buff.m:1-33."""
    ctx = get_context(device)
    mag = np.asarray(TF_mag, dtype=np.float64)
    if mag.ndim == 1:
        mag = mag[:, None]
    mag = _f64(mag)
    ph = None
    if TF_phase is not None:
        ph = _f64(np.asarray(TF_phase, dtype=np.float64).reshape(mag.shape))
    win = _f64(np.asarray(win, dtype=np.float64).ravel())
    out = np.empty((int(sz), mag.shape[1]), order="F")
    check(ctx._lib.snmfnat_synth_ifft_buff(ctx._h, _dptr(mag), _dptr(ph), mag.shape[0], mag.shape[1], int(sz), int(fftlen),
                                           _dptr(win), float(preemph), int(DCbin_back), float(pow_), _dptr(out)))
    return out


def blk_sparse(X, D, r_blk, l, p: dict, *, device: int = 0):
    """[Q, r_blk_out] = blk_sparse(X, D, r_blk, l, p)      src/blk_sparse.m:1-37."""
    ctx = get_context(device)
    X = _f64(np.asarray(X, dtype=np.float64).ravel())
    D = _f64(np.asarray(D, dtype=np.float64).ravel())
    rb = _f64(r_blk)
    ps = params_struct(p)
    if rb.shape != (X.size, ps.P_len_l):
        raise ValueError("r_blk must be K x P_len_l")
    Q = np.empty(X.size)
    rout = np.empty_like(rb, order="F")
    check(ctx._lib.snmfnat_blk_sparse(ctx._h, _dptr(X), _dptr(D), _dptr(rb), X.size, int(l), C.byref(ps), _dptr(Q),
                                      _dptr(rout)))
    return Q, rout


# ----------------------------------------------------------------------------------------------------------------
# L2 per-hop entry: init_buff / bnmf_sep_event_RT_IS16 with a device-resident g
# ----------------------------------------------------------------------------------------------------------------
class StreamState:
    """The struct g of src/init_buff.m:17-62, resident on the device.  g["B_DFT_d"], g["Ad_blk"], ... read the fields
    (MATLAB layouts); assignment writes them back (e.g. the B_D_u.mat carry-over of src/NTF_sep_event_RT.m:28-38)."""

    _SHAPES = {"B_DFT_d": ("F", "R_d"), "B_Mel_d": ("F", "R_d"), "B_DFT_x": ("F", "R_x"), "B_Mel_x": ("F", "R_x"),
               "Ad_blk": ("R_a", "m_a"), "lambda_d_blk": ("F", "m_a"), "r_blk": ("F", "P_len_l"),
               "lambda_dav": ("F",), "Xm_tilde": ("F",), "Ym": ("F",), "Yp": ("F",), "A": ("R",), "Q": ("F",),
               "G": ("F",), "Xm_hat": ("F",), "Dm_hat": ("F",), "update_switch": (1,), "stats": (5,)}

    def __init__(self, ctx: Context, handle, ps: Params, n1=None):
        self.ctx, self._h, self._ps = ctx, handle, ps
        self._dims = {"F": ps.fftlength // 2 + 1, "R_x": ps.R_x, "R_d": ps.R_d, "R": ps.R_x + ps.R_d, "R_a": ps.R_a,
                      "m_a": ps.m_a, "P_len_l": ps.P_len_l}
        # B_sep_mode = 'Mel': the Mel dictionaries have n1 rows (the noise history g.lambda_d_blk stays in the DFT domain)
        self._mel_rows = int(n1) if (ps.B_sep_mode == _lib.SEP_MEL and n1) else None

    def _shape(self, name):
        if name not in self._SHAPES:
            raise KeyError(name)
        shp = tuple(self._dims.get(x, x) for x in self._SHAPES[name])
        if self._mel_rows and name in ("B_Mel_d", "B_Mel_x"):
            shp = (self._mel_rows,) + shp[1:]
        return shp

    def __getitem__(self, name):
        shp = self._shape(name)
        buf = np.empty(shp, dtype=np.float64, order="F")
        check(self.ctx._lib.snmfnat_stream_get(self._h, name.encode(), _dptr(buf), buf.size))
        return buf

    def __setitem__(self, name, value):
        shp = self._shape(name)
        buf = _f64(np.asarray(value, dtype=np.float64).reshape(shp))
        check(self.ctx._lib.snmfnat_stream_set(self._h, name.encode(), _dptr(buf), buf.size))

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.snmfnat_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def init_buff(B_Mel_x, B_Mel_d, B_DFT_x, B_DFT_d, p: dict, *, Ad_blk_init, A_d_init=None, device: int = 0) -> StreamState:
    """g = init_buff(B_Mel_x, B_Mel_d, B_DFT_x, B_DFT_d, p)   src/init_buff.m:1-62.  The two rand() draws at :37-38
    are arguments (Ad_blk_init R_a x m_a, A_d_init R_d x 1)."""
    ctx = get_context(device)
    ps = params_struct(p)
    Bmx, Bmd, Bx, Bd = _f64(B_Mel_x), _f64(B_Mel_d), _f64(B_DFT_x), _f64(B_DFT_d)
    win_s = _f64(np.asarray(p.get("win_STFT", sqrt_hann_periodic(ps.framelength))).ravel())
    win_i = _f64(np.asarray(p.get("win_ISTFT", sqrt_hann_periodic(ps.framelength))).ravel())
    ad = _f64(np.asarray(Ad_blk_init, dtype=np.float64).reshape(ps.R_a, ps.m_a)) if Ad_blk_init is not None else None
    a_d = _f64(np.asarray(A_d_init, dtype=np.float64).ravel()) if A_d_init is not None else None
    h = C.c_void_p()
    check(ctx._lib.snmfnat_stream_create(ctx._h, C.byref(ps), _dptr(win_s), _dptr(win_i), _dptr(Bmx), _dptr(Bmd),
                                         Bmd.shape[0], _dptr(Bx), _dptr(Bd), Bx.shape[0], _dptr(ad), _dptr(a_d),
                                         C.byref(h)))
    return StreamState(ctx, h, ps, n1=Bmd.shape[0])


def bnmf_sep_event_RT_IS16(y, l, g: StreamState, p: dict, *, h_init, nargout: int = 3):
    """[x_hat_i, d_hat_i, x_tilde, g] = bnmf_sep_event_RT_IS16(y, l, g, p)    src/bnmf_sep_event_RT_IS16.m:1-423.
    y: framelength samples (ch = 1); l: 1-based hop index; h_init: rand(R_x+R_d, 1) after rand('seed', p.random_seed).
    With nargout <= 3 and callers that discard x_hat_i / d_hat_i (filewise_run_IS16.m:142) pass nargout=1 to skip the
    two extra ISTFTs: they are then returned as None."""
    ps = g._ps
    y = _f64(np.asarray(y, dtype=np.float64).ravel())
    if y.size != ps.framelength:
        raise ValueError("y must hold p.framelength samples")
    h0 = _f64(np.asarray(h_init, dtype=np.float64).ravel())
    xt = np.empty(ps.framelength)
    xh = dh = None
    if nargout >= 2:
        xh = np.empty((ps.EVENT_NUM, ps.framelength), order="C")
        dh = np.empty((ps.NOISE_NUM, ps.framelength), order="C")
    check(g.ctx._lib.snmfnat_stream_step(g._h, _dptr(y), int(l), _dptr(h0), _dptr(xt), _dptr(xh), _dptr(dh)))
    if xh is not None:
        xh = xh.reshape(ps.EVENT_NUM, 1, ps.framelength)      # :373-380 (3-D for compatibility with the NTF functions)
        dh = dh.reshape(1, ps.NOISE_NUM, ps.framelength)
    return xh, dh, xt, g


# ----------------------------------------------------------------------------------------------------------------
# offline dictionary training (run_basis_train.m:80-91,112-116) -- frame-sharded, tf32 tensor cores
# ----------------------------------------------------------------------------------------------------------------
class Train:
    """Frame shard of one `sparse_nmf(TF_mag, p)` call with W and H both updated (run_basis_train.m:84-88),
    resident on one GPU.  V is F x T_local, init_w F x K, init_h K x T_local (MATLAB layouts; float32 on the device)."""

    def __init__(self, ctx: Context, F: int, K: int, T_local: int, sparsity: float):
        self._lib = _lib.load()
        self.ctx = ctx
        self.F, self.K, self.T = int(F), int(K), int(T_local)
        h = C.c_void_p()
        check(self._lib.snmfnat_train_create(ctx._h, self.F, self.K, self.T, float(sparsity), 1, C.byref(h)))
        self._h = h
        ldv, kp = C.c_int(), C.c_int()
        check(self._lib.snmfnat_train_get_layout(self._h, C.byref(ldv), C.byref(kp)))
        self.ldv, self.Kp = ldv.value, kp.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.snmfnat_train_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def attach_nccl(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(self._lib.snmfnat_train_attach_nccl(self._h, C.cast(buf, C.c_void_p), int(rank), int(world)))

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().snmfnat_train_nccl_unique_id(C.cast(buf, C.c_void_p)))
        return buf.raw

    def set_data(self, V=None, init_w=None, init_h=None):
        """Host arrays in MATLAB shapes (F x T, F x K, K x T)."""
        def f32(a, shape):
            if a is None:
                return None, None
            a = np.asfortranarray(np.asarray(a, dtype=np.float32))
            assert a.shape == shape, (a.shape, shape)
            return a, a.ctypes.data_as(C.c_void_p)
        v, vp = f32(V, (self.F, self.T))
        w, wp = f32(init_w, (self.F, self.K))
        h, hp = f32(init_h, (self.K, self.T))
        check(self._lib.snmfnat_train_set_data(self._h, vp, 0, wp, hp, 0))

    def dev_ptr(self, which: str) -> int:
        return int(self._lib.snmfnat_train_dev_ptr(self._h, which.encode()) or 0)

    def commit_v(self):
        """After filling dev_ptr("V") in place: rebuild the bin-major copy of V."""
        check(self._lib.snmfnat_train_commit_v(self._h))

    def reset(self):
        check(self._lib.snmfnat_train_reset(self._h))

    def iterate(self, n: int, want_cost: bool = False):
        if not want_cost:
            check(self._lib.snmfnat_train_iterate(self._h, int(n), None, None))
            return None
        div = np.zeros(n)
        cost = np.zeros(n)
        check(self._lib.snmfnat_train_iterate(self._h, int(n), _dptr(div), _dptr(cost)))
        return dict(div=div, cost=cost)

    def run(self, max_iter: int, conv_eps: float):
        div = np.zeros(max_iter)
        cost = np.zeros(max_iter)
        its = C.c_int()
        check(self._lib.snmfnat_train_run(self._h, int(max_iter), float(conv_eps), _dptr(div), _dptr(cost), C.byref(its)))
        n = its.value
        return dict(div=div[:n], cost=cost[:n], iters=n)

    def get_w(self) -> np.ndarray:
        w = np.zeros((self.F, self.K), dtype=np.float32, order="F")
        check(self._lib.snmfnat_train_get_w(self._h, w.ctypes.data_as(C.POINTER(C.c_float))))
        return w

    def get_acc(self):
        g = np.zeros((self.F, self.K), dtype=np.float32, order="F")
        hs = np.zeros(self.K, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        check(self._lib.snmfnat_train_get_acc(self._h, g.ctypes.data_as(fp), hs.ctypes.data_as(fp)))
        return g, hs

    def get_h(self, t0: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.T - t0 if count is None else count
        h = np.zeros((self.K, count), dtype=np.float32, order="F")
        check(self._lib.snmfnat_train_get_h(self._h, h.ctypes.data_as(C.POINTER(C.c_float)), int(t0), int(count)))
        return h


def basis_train_core(TF_pow, R: int, sample_idx, p: dict, *, h_init, device: int = 0, comm=None):
    """Numeric core of run_basis_train.m:80-91,112-116 on the GPU: exemplar init B_init = TF_mag(:, idx) (:81-83),
    sparse_nmf with W and H updated (:84-88), column normalisation + 1e-9 (:112-114).  `sample_idx` replaces
    randsample, `h_init` is the rand(R, T) of sparse_nmf.m:134.  Returns (B, A, objective)."""
    TF_pow = np.asarray(TF_pow)
    F, T = TF_pow.shape
    if str(p.get("cf", "kl")) != "kl":
        raise SnmfnatError(-4, "the tensor-core training path implements cf='kl' (use sparse_nmf for other divergences)")
    tr = Train(get_context(device), F, R, T, float(p["sparsity"]))
    try:
        tr.set_data(TF_pow, TF_pow[:, np.asarray(sample_idx)], h_init)
        if p.get("cost_check", 1):
            obj = tr.run(int(p["max_iter"]), float(p["conv_eps"]))
        else:
            tr.iterate(int(p["max_iter"]))
            obj = dict(div=np.zeros(0), cost=np.zeros(0), iters=int(p["max_iter"]))
        w = tr.get_w().astype(np.float64)
        h = tr.get_h().astype(np.float64)
    finally:
        tr.close()
    wn = np.sqrt(np.sum(w ** 2, axis=0))
    return w / wn + 1e-9, h, obj
