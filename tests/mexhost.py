"""ctypes driver of mex/mex_host.cpp: builds mxArrays from NumPy / dicts, calls a gateway's mexFunction, reads the
outputs back.  The gateways under mex/_build/*.so are the product's MEX sources linked against this fake host and
libsnmfnat.so (`make -C mex host`)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
BUILD = ROOT / "mex" / "_build"
DOUBLE, UINT64, LOGICAL, CHAR, STRUCT, CELL, INT16 = 1, 2, 3, 4, 5, 6, 7   # mxClassID of mex/mex_shim.h


def build():
    r = subprocess.run(["make", "-C", str(ROOT / "mex"), "host"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)


class Host:
    def __init__(self):
        if not (BUILD / "libmexhost.so").exists():
            build()
        self.lib = L = C.CDLL(str(BUILD / "libmexhost.so"), mode=C.RTLD_GLOBAL)
        vp, sz = C.c_void_p, C.c_size_t
        for name, res, args in [
            ("mxCreateDoubleMatrix", vp, [sz, sz, C.c_int]), ("mxCreateDoubleScalar", vp, [C.c_double]),
            ("mxCreateNumericArray", vp, [sz, C.POINTER(sz), C.c_int, C.c_int]), ("mxCreateString", vp, [C.c_char_p]),
            ("mxCreateLogicalMatrix", vp, [sz, sz]), ("mxCreateCellMatrix", vp, [sz, sz]),
            ("mxCreateStructMatrix", vp, [sz, sz, C.c_int, C.POINTER(C.c_char_p)]), ("mxAddField", C.c_int, [vp, C.c_char_p]),
            ("mxSetField", None, [vp, sz, C.c_char_p, vp]), ("mxSetCell", None, [vp, sz, vp]), ("mxGetField", vp, [vp, sz, C.c_char_p]),
            ("mxGetData", vp, [vp]), ("mxGetNumberOfDimensions", sz, [vp]), ("mxGetDimensions", C.POINTER(sz), [vp]),
            ("mxGetNumberOfElements", sz, [vp]), ("mxDestroyArray", None, [vp]), ("mxDuplicateArray", vp, [vp]),
            ("mexhost_call", C.c_int, [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_char_p, C.c_int]),
            ("mexhost_set_rand_stream", None, [C.POINTER(C.c_double), sz]), ("mexhost_shutdown", None, []),
            ("mexhost_live_arrays", C.c_longlong, []), ("mexhost_class", C.c_int, [vp]),
            ("mexhost_num_fields", C.c_int, [vp]), ("mexhost_field_name", C.c_char_p, [vp, C.c_int]),
        ]:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self._gw = {}

    # ---- NumPy / Python -> mxArray
    def mx(self, v):
        L = self.lib
        if isinstance(v, dict):
            names = list(v)
            arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
            s = L.mxCreateStructMatrix(1, 1, len(names), arr)
            for n in names:
                L.mxSetField(s, 0, n.encode(), self.mx(v[n]))
            return s
        if isinstance(v, str):
            return L.mxCreateString(v.encode())
        if isinstance(v, (list, tuple)) and v and isinstance(v[0], str):
            c = L.mxCreateCellMatrix(1, len(v))
            for i, x in enumerate(v):
                L.mxSetCell(c, i, self.mx(x))
            return c
        a = np.asarray(v)
        if a.dtype == np.bool_:
            a2 = np.atleast_2d(a) if a.ndim < 2 else a
            if a.ndim == 1:
                a2 = a.reshape(-1, 1)
            m = L.mxCreateLogicalMatrix(a2.shape[0], a2.shape[1])
            C.memmove(L.mxGetData(m), np.asfortranarray(a2.astype(np.uint8)).ctypes.data, a2.size)
            return m
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 0:
            return L.mxCreateDoubleScalar(float(a))
        if a.ndim == 1:
            a = a.reshape(1, -1)            # MATLAB row vector
        dims = (C.c_size_t * a.ndim)(*a.shape)
        m = L.mxCreateNumericArray(a.ndim, dims, DOUBLE, 0)
        f = np.asfortranarray(a)
        C.memmove(L.mxGetData(m), f.ctypes.data, f.nbytes)
        return m

    # ---- mxArray -> NumPy / dict
    def py(self, m):
        L = self.lib
        if not m:
            return None
        cls = L.mexhost_class(m)
        nd = L.mxGetNumberOfDimensions(m)
        dims = [L.mxGetDimensions(m)[i] for i in range(nd)]
        n = L.mxGetNumberOfElements(m)
        if cls == STRUCT:
            return {L.mexhost_field_name(m, i).decode(): self.py(L.mxGetField(m, 0, L.mexhost_field_name(m, i)))
                    for i in range(L.mexhost_num_fields(m))}
        if cls == CHAR:
            return C.string_at(L.mxGetData(m), n).decode()
        dt = {DOUBLE: np.float64, UINT64: np.uint64, LOGICAL: np.uint8, INT16: np.int16}[cls]
        buf = np.frombuffer(C.string_at(L.mxGetData(m), n * np.dtype(dt).itemsize), dtype=dt)
        return buf.reshape(dims, order="F").copy()

    def gateway(self, name):
        if name not in self._gw:
            self._gw[name] = C.CDLL(str(BUILD / f"{name}.so"))
        return self._gw[name]

    def call(self, name, nlhs, *args, keep=False):
        """Run mexFunction of gateway `name`.  args: Python values or raw mxArray handles (int).  Returns the outputs
        converted to Python (or the raw handles with keep=True: the caller destroys them)."""
        L = self.lib
        fn = C.cast(self.gateway(name).mexFunction, C.c_void_p)
        made = []
        ins = []
        for a in args:
            if isinstance(a, int):
                ins.append(a)
            else:
                h = self.mx(a)
                made.append(h)
                ins.append(h)
        prhs = (C.c_void_p * max(len(ins), 1))(*ins)
        plhs = (C.c_void_p * max(nlhs, 1))()
        err = C.create_string_buffer(2048)
        rc = L.mexhost_call(fn, nlhs, plhs, len(ins), prhs, err, 2048)
        for h in made:
            L.mxDestroyArray(h)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        outs = [plhs[i] for i in range(nlhs)]
        if keep:
            return outs
        res = [self.py(o) for o in outs]
        for o in outs:
            if o:
                L.mxDestroyArray(o)
        return res

    def set_rand_stream(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.lib.mexhost_set_rand_stream(v.ctypes.data_as(C.POINTER(C.c_double)), v.size)


def matlab_p(p: dict) -> dict:
    """api.default_p()-style dict -> the fields of the reference's global p as MATLAB would hold them."""
    q = {}
    for k, v in p.items():
        if k in ("win_STFT", "win_ISTFT"):
            q[k] = np.asarray(v, dtype=np.float64).reshape(-1, 1)
        elif k in ("EVENT_RANK", "NOISE_RANK"):
            q[k] = np.asarray(v, dtype=np.float64)
        elif k in ("cf", "ENHANCE_METHOD", "B_sep_mode"):
            q[k] = v if isinstance(v, str) else {"cf": "kl", "ENHANCE_METHOD": "MMSE", "B_sep_mode": "DFT"}[k]
        elif k == "beta_div":
            continue
        else:
            q[k] = float(v) if np.isscalar(v) else np.asarray(v, dtype=np.float64)
    return q
