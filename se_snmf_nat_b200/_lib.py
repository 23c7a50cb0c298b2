"""ctypes binding of libsnmfnat.so (include/snmfnat.h).

The library is the product; this module only loads it and declares the
signatures.  There is no fallback: if the shared object is missing or a CUDA
device is not usable the calls raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libsnmfnat.so"

MAX_CLASSES = 8

CF_IS, CF_KL, CF_ED, CF_BETA = 0, 1, 2, 3
ENH_MMSE, ENH_WIENER = 0, 1
SEP_DFT, SEP_MEL = 0, 1

ERR_NAMES = {0: "OK", -1: "EINVAL", -2: "ENODEVICE", -3: "ECUDA", -4: "EUNSUPPORTED", -5: "ENUMERIC", -6: "ENOMEM"}


class SnmfnatError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsnmfnat {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Params(C.Structure):
    """struct snmfnat_params (include/snmfnat.h)."""
    _fields_ = [
        ("fs", C.c_int32), ("framelength", C.c_int32), ("frameshift", C.c_int32), ("fftlength", C.c_int32),
        ("delay", C.c_int32),
        ("blk_len_sep", C.c_int32), ("blk_hop_sep", C.c_int32), ("Splice", C.c_int32),
        ("EVENT_NUM", C.c_int32), ("NOISE_NUM", C.c_int32),
        ("EVENT_RANK", C.c_int32 * MAX_CLASSES), ("NOISE_RANK", C.c_int32 * MAX_CLASSES),
        ("R_x", C.c_int32), ("R_d", C.c_int32), ("R_a", C.c_int32), ("m_a", C.c_int32),
        ("init_N_len", C.c_int32), ("adapt_train_N", C.c_int32),
        ("blk_sparse", C.c_int32), ("P_len_k", C.c_int32), ("P_len_l", C.c_int32), ("blk_gap", C.c_int32),
        ("DCbin", C.c_int32), ("DCbin_back", C.c_int32), ("F_order", C.c_int32),
        ("B_sep_mode", C.c_int32), ("MelConv", C.c_int32),
        ("cf", C.c_int32), ("max_iter", C.c_int32), ("cost_check", C.c_int32),
        ("basis_update_N", C.c_int32), ("basis_update_E", C.c_int32),
        ("ENHANCE_METHOD", C.c_int32),
        ("reserved_i", C.c_int32 * 7),
        ("overlapscale", C.c_double), ("pow", C.c_double), ("nonzerofloor", C.c_double),
        ("overlap_m_a", C.c_double), ("Ar_up", C.c_double), ("alpha_p", C.c_double), ("preemph", C.c_double),
        ("beta_div", C.c_double),
        ("sparsity", C.c_double), ("conv_eps", C.c_double),
        ("alpha_eta", C.c_double), ("alpha_d", C.c_double), ("beta", C.c_double), ("beta_max", C.c_double),
        ("sparsity_mdi", C.c_double), ("conv_eps_mdi", C.c_double),
        ("reserved_d", C.c_double * 6),
    ]


class NmfOpts(C.Structure):
    """struct snmfnat_nmf_opts."""
    _fields_ = [
        ("max_iter", C.c_int32), ("cf", C.c_int32), ("cost_check", C.c_int32),
        ("sparsity_rows", C.c_int32), ("sparsity_cols", C.c_int32), ("precision", C.c_int32),
        ("beta_div", C.c_double), ("conv_eps", C.c_double),
    ]


class BatchStats(C.Structure):
    """struct snmfnat_batch_stats."""
    _fields_ = [
        ("hops", C.c_int64), ("h_iters", C.c_int64), ("w_iters", C.c_int64), ("gated_hops", C.c_int64),
        ("w_solves", C.c_int64), ("w_atoms", C.c_int64), ("flops", C.c_double), ("launches", C.c_int64),
        ("reserved", C.c_int64 * 4),
    ]


_P = C.POINTER
_dp = _P(C.c_double)
_fp = _P(C.c_float)
_u8p = _P(C.c_uint8)
_vp = C.c_void_p
_i16pp = _P(_P(C.c_int16))

# name -> (restype, argtypes); every symbol include/snmfnat.h declares
SIGNATURES = {
    "snmfnat_version": (C.c_int, []),
    "snmfnat_ctx_create": (C.c_int, [C.c_int, _P(_vp)]),
    "snmfnat_ctx_destroy": (C.c_int, [_vp]),
    "snmfnat_ctx_sync": (C.c_int, [_vp]),
    "snmfnat_ctx_cuda_stream": (_vp, [_vp]),
    "snmfnat_last_error": (C.c_char_p, [_vp]),
    "snmfnat_ctx_launch_count": (C.c_int64, [_vp]),
    "snmfnat_params_default": (None, [_P(Params)]),
    "snmfnat_sparse_nmf": (C.c_int, [_vp, _dp, C.c_int, C.c_int, C.c_int, _P(NmfOpts), _dp, _dp, _dp, _u8p, _u8p,
                                     _dp, _dp, _dp, _dp, _P(C.c_int)]),
    "snmfnat_snmf_mdi": (C.c_int, [_vp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _P(NmfOpts), _dp, _dp, _dp,
                                   _u8p, _u8p, _dp, _dp, _dp, _dp, _P(C.c_int)]),
    "snmfnat_dnmf_adapt": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _P(NmfOpts), _dp, _dp,
                                     _dp]),
    "snmfnat_stft_fft": (C.c_int, [_vp, _dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double, _dp,
                                   _dp]),
    "snmfnat_synth_ifft_buff": (C.c_int, [_vp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_int,
                                          C.c_double, _dp]),
    "snmfnat_blk_sparse": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int, _P(Params), _dp, _dp]),
    "snmfnat_stream_create": (C.c_int, [_vp, _P(Params), _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, C.c_int, _dp, _dp,
                                        _P(_vp)]),
    "snmfnat_stream_destroy": (C.c_int, [_vp]),
    "snmfnat_stream_step": (C.c_int, [_vp, _dp, C.c_int, _dp, _dp, _dp, _dp]),
    "snmfnat_stream_get": (C.c_int, [_vp, C.c_char_p, _dp, C.c_int64]),
    "snmfnat_stream_set": (C.c_int, [_vp, C.c_char_p, _dp, C.c_int64]),
    "snmfnat_batch_create": (C.c_int, [_vp, _P(Params), _dp, _dp, _dp, _dp, C.c_int, C.c_int, _P(C.c_int64),
                                       _P(C.c_int32), _dp, _dp, C.c_int64, _P(_vp)]),
    "snmfnat_batch_destroy": (C.c_int, [_vp]),
    "snmfnat_batch_upload": (C.c_int, [_vp, _i16pp]),
    "snmfnat_batch_upload_packed": (C.c_int, [_vp, _P(C.c_int16)]),
    "snmfnat_batch_run": (C.c_int, [_vp]),
    "snmfnat_batch_download": (C.c_int, [_vp, _i16pp]),
    "snmfnat_batch_download_packed": (C.c_int, [_vp, _P(C.c_int16)]),
    "snmfnat_batch_out_len": (C.c_int64, [_vp, C.c_int]),
    "snmfnat_batch_total_hops": (C.c_int64, [_vp]),
    "snmfnat_batch_get_stats": (C.c_int, [_vp, _P(BatchStats)]),
    "snmfnat_batch_enable_trace": (C.c_int, [_vp, C.c_int]),
    "snmfnat_batch_get_trace": (C.c_int, [_vp, C.c_int, C.c_char_p, _dp, C.c_int64]),
    "snmfnat_batch_set_profile": (C.c_int, [_vp, C.c_int]),
    "snmfnat_batch_set_groups": (C.c_int, [_vp, C.c_int]),
    "snmfnat_batch_set_mel": (C.c_int, [_vp, _dp, _dp, C.c_int, _dp]),
    "snmfnat_tf_features": (C.c_int, [_vp, _dp, C.c_int, C.c_int64, C.c_double, C.c_double, _dp, C.c_int, _dp]),
    "snmfnat_gist_ntf": (C.c_int, [_vp, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_int,
                                  C.c_double, C.c_int, _dp, _dp, _dp, _P(C.c_int)]),
    "snmfnat_mel_matrix": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp]),
    "snmfnat_batch_get_profile": (C.c_int, [_vp, _dp, _P(C.c_int64)]),
    "snmfnat_batch_get_noise_basis": (C.c_int, [_vp, C.c_int, _dp]),
    "snmfnat_enhance_batch": (C.c_int, [_vp, _P(Params), _dp, _dp, _dp, _dp, C.c_int, C.c_int, _i16pp,
                                        _P(C.c_int64), _P(C.c_int32), _dp, _dp, C.c_int64, _i16pp]),
    "snmfnat_enhance_batch_multi": (C.c_int, [_P(C.c_int), C.c_int, _P(Params), _dp, _dp, _dp, _dp, C.c_int, C.c_int, _i16pp,
                                              _P(C.c_int64), _P(C.c_int32), _dp, _dp, C.c_int64, _i16pp]),
    "snmfnat_train_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_int, _P(_vp)]),
    "snmfnat_train_destroy": (C.c_int, [_vp]),
    "snmfnat_train_nccl_unique_id": (C.c_int, [_vp]),
    "snmfnat_train_attach_nccl": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "snmfnat_train_set_data": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "snmfnat_train_dev_ptr": (_vp, [_vp, C.c_char_p]),
    "snmfnat_train_get_layout": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int)]),
    "snmfnat_train_commit_v": (C.c_int, [_vp]),
    "snmfnat_train_reset": (C.c_int, [_vp]),
    "snmfnat_train_iterate": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "snmfnat_train_run": (C.c_int, [_vp, C.c_int, C.c_double, _dp, _dp, _P(C.c_int)]),
    "snmfnat_train_get_acc": (C.c_int, [_vp, _fp, _fp]),
    "snmfnat_train_get_w": (C.c_int, [_vp, _fp]),
    "snmfnat_train_get_h": (C.c_int, [_vp, _fp, C.c_int64, C.c_int64]),
}

_lib = None


def load() -> C.CDLL:
    """Load libsnmfnat.so (built in-tree by ``python -m se_snmf_nat_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m se_snmf_nat_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().snmfnat_last_error(None)
        raise SnmfnatError(rc, msg.decode() if msg else "")
