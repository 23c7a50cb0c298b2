"""Reads every settings file the reference ships (settings/*.m, settings/bak_IS16_results/*.m) and writes the fields the
hot path consumes to tests/golden/reference_settings.json, so that the parity tests can run the reference's OWN parameter
sets on a box that has no copy of the reference.

    python tests/golden/make_settings_fixture.py [/root/reference]

The files are flat lists of `p.name = expression;` assignments; the expressions use a handful of MATLAB functions
(floor, ceil, log2, round, sqrt, hann, num2str) and earlier fields of p, so they are evaluated in file order by a small
translator instead of being copied.  Fields a file does not set stay absent: the library (snmfnat_params_default) and the
oracle (default_params) fill them with the shipped values, exactly as a caller who only overrides what differs would.
"""
import json
import math
import re
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

# what struct snmfnat_params / oracle.default_params know about (numbers), plus the three string switches
FIELDS = ["fs", "framelength", "frameshift", "fftlength", "delay", "blk_len_sep", "blk_hop_sep", "Splice", "EVENT_NUM",
          "NOISE_NUM", "EVENT_RANK", "NOISE_RANK", "R_x", "R_d", "R_a", "m_a", "init_N_len", "adapt_train_N", "blk_sparse",
          "P_len_k", "P_len_l", "blk_gap", "DCbin", "DCbin_back", "F_order", "MelConv", "max_iter", "cost_check",
          "basis_update_N", "basis_update_E", "overlapscale", "pow", "nonzerofloor", "overlap_m_a", "Ar_up", "alpha_p",
          "preemph", "sparsity", "conv_eps", "alpha_eta", "alpha_d", "beta", "beta_max", "sparsity_mdi", "conv_eps_mdi",
          "random_seed", "useGPU"]
STRINGS = ["NMF_algorithm", "B_sep_mode", "cf", "ENHANCE_METHOD"]


def _hann_periodic(n, _flag=None):
    n = int(n)
    return 0.5 * (1 - np.cos(2 * np.pi * np.arange(n) / n))


def _hanning(n):
    n = int(n)
    return 0.5 * (1 - np.cos(2 * np.pi * np.arange(1, n + 1) / (n + 1)))


ENV = dict(floor=math.floor, ceil=math.ceil, log2=math.log2, round=lambda x: int(math.floor(x + 0.5)), sqrt=np.sqrt,
           hann=_hann_periodic, hanning=_hanning, num2str=str, sum=np.sum, inf=float("inf"))


def parse(path):
    p = {}
    for raw in open(path, encoding="latin-1"):
        line = raw.split("%")[0].strip()
        m = re.match(r"^p\.(\w+)(\(\d+\))?\s*=\s*(.+?);?\s*$", line)
        if not m:
            continue
        name, idx, rhs = m.groups()
        py = re.sub(r"p\.(\w+)", r"p['\1']", rhs.rstrip(";").strip()).replace("^", "**")
        py = re.sub(r"\[([^\]]*)\]",
                    lambda mm: "[" + ",".join(x for x in re.split(r"[,\s]+", mm.group(1).strip()) if x) + "]", py)
        try:
            v = eval(py, dict(ENV), dict(p=p))   # noqa: S307 - our own translation of a settings line
        except Exception:                        # output names built with strcat etc.: not parameters of the path
            continue
        if idx:                                  # p.EVENT_RANK(2) = 21
            i = int(idx[1:-1]) - 1
            lst = list(p.get(name, []))
            lst += [0] * (i + 1 - len(lst))
            lst[i] = v
            p[name] = lst
        else:
            p[name] = v
    return p


def main():
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    out = {}
    for f in sorted((ref / "settings").rglob("*.m")):
        p = parse(f)
        rec = {}
        for k in FIELDS + STRINGS:
            if k in p:
                v = p[k]
                if isinstance(v, (list, tuple)):
                    v = [int(x) for x in v]
                elif isinstance(v, (np.floating, float)):
                    v = float(v)
                    if v == int(v) and k not in ("overlapscale", "pow", "nonzerofloor", "overlap_m_a", "Ar_up", "alpha_p",
                                                 "preemph", "sparsity", "conv_eps", "alpha_eta", "alpha_d", "beta",
                                                 "beta_max", "sparsity_mdi", "conv_eps_mdi"):
                        v = int(v)
                rec[k] = v
        # the windows are sqrt(hann(framelength,'periodic')) in every file; record that it is so instead of 640 numbers
        w = p.get("win_STFT")
        rec["_win_is_sqrt_hann_periodic"] = bool(
            w is not None and np.allclose(w, np.sqrt(_hann_periodic(p["framelength"]))) and
            np.allclose(p.get("win_ISTFT"), w))
        out[str(f.relative_to(ref))] = rec
    (HERE / "reference_settings.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    for k, v in out.items():
        print(k, len(v), "fields")


if __name__ == "__main__":
    main()
