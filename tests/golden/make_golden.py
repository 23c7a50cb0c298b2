"""Generates the committed fixtures under tests/golden/ from the reference's shipped data.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What it writes
  bases.npz        B_DFT_x / B_DFT_d / B_Mel_x / B_Mel_d of basis/*/TASLP_Splice0-SNMF_p2_DD0/R_100.mat (MAT v5)
  wavs.npz         int16 samples (44-byte header stripped) of the reference's wav/ inputs and its two shipped
                   outputs (*_out_v3.9_18.wav), the only end-to-end golden vectors the reference holds
  rng_seed1.npz    the stand-ins for MATLAB's RNG draws (h_init = rand(200,1) after rand('seed',1), then in the order of
                   init_buff.m:37-38 A_d = rand(100,1), Ad_blk = rand(50,100)); see oracle.snmf_oracle.default_rng_inputs
  rng_seed1.mat    the same stream as MAT v5 for tests/golden/ref_shadow/rand.m (make_ref_vectors.m)
  M03_oracle.npz   the float64 oracle's own result on M03 (output, per-hop iteration counts / gates, a few
                   activations, final noise basis) -- pins the oracle against regressions
"""
import sys
from pathlib import Path

import numpy as np
import scipy.io

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import snmf_oracle as O  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def main():
    bx = scipy.io.loadmat(REF / "basis/Clean_train_TIMIT_test/TASLP_Splice0-SNMF_p2_DD0/R_100.mat")
    bd = scipy.io.loadmat(REF / "basis/CHiME3_bgn_ch6/TASLP_Splice0-SNMF_p2_DD0/R_100.mat")
    np.savez_compressed(OUT / "bases.npz", B_DFT_x=bx["B_DFT_sub"], B_DFT_d=bd["B_DFT_sub"],
                        B_Mel_x=bx["B_Mel_sub"], B_Mel_d=bd["B_Mel_sub"])
    wavs = {}
    for name, key in [("M03_423C0213_STR.CH6", "M03_in"), ("M03_423C0213_STR.CH6_out_v3.9_18", "M03_ref_out"),
                      ("M04_423C020A_STR.CH6", "M04_in"), ("LM_in", "LM_in"), ("LM_in_out_v3.9_18", "LM_ref_out")]:
        wavs[key] = O.read_wav_pcm(REF / "wav" / f"{name}.wav")
    np.savez_compressed(OUT / "wavs.npz", **wavs)
    p = O.default_params()
    h_init, Ad, A_d = O.default_rng_inputs(p, with_A_d=True)
    np.savez_compressed(OUT / "rng_seed1.npz", h_init=h_init, Ad_blk=Ad, A_d=A_d)
    # the same draws for tests/golden/ref_shadow/rand.m (Octave / MATLAB): init_buff.m:37-38 order, column-major
    scipy.io.savemat(OUT / "rng_seed1.mat", {"stream": np.concatenate([A_d, Ad.T.ravel()])[:, None], "h_init": h_init[:, None]},
                     format="5")
    tr = []
    out, g = O.enhance_utterance(wavs["M03_in"], p, bx["B_DFT_sub"], bd["B_DFT_sub"], h_init=h_init, Ad_blk_init=Ad,
                                 trace=tr)
    np.savez_compressed(
        OUT / "M03_oracle.npz", out=out,
        h_iters=np.array([t["h_iters"] for t in tr]), gated=np.array([t["gated"] for t in tr]),
        R_a_up=np.array([t["R_a_up"] for t in tr]), w_iters=np.array([t["w_iters"] for t in tr]),
        A=np.stack([t["A"] for t in tr]), Xm_tilde=np.stack([t["Xm_tilde"] for t in tr]),
        B_DFT_d_final=g.B_DFT_d)
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
