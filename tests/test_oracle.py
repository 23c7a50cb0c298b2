"""CPU tests of the float64 oracle: it is pinned against the reference's own golden wav pairs (the only
known-answer vectors the reference ships) and against structural invariants."""
import numpy as np
import pytest

from oracle import snmf_oracle as O
from conftest import snr_db


def test_oracle_matches_reference_golden_wav_M03(bases, wavs, rng_inputs, m03_oracle):
    p = O.default_params()
    h_init, Ad = rng_inputs
    tr = []
    out, g = O.enhance_utterance(wavs["M03_in"], p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad,
                                 trace=tr)
    gold = wavs["M03_ref_out"]
    # exact output length: (floor(N/160)+1)*160   (filewise_run_IS16.m:146,162-165)
    assert len(out) == len(gold) == (len(wavs["M03_in"]) // 160 + 1) * 160 == 55040
    # coarse end-to-end pin: the shipped output was made with MATLAB RNG streams we cannot regenerate; four other RNG
    # streams give 21.7 .. 24.1 dB against it, so 21 dB is as tight as this pin can be (see the negative controls below)
    assert snr_db(gold, out) > 21.0
    # the oracle's own committed result: bit-exact regression pin
    assert np.array_equal(out, m03_oracle["out"])
    assert np.array_equal([t["h_iters"] for t in tr], m03_oracle["h_iters"])
    assert np.array_equal([t["R_a_up"] for t in tr], m03_oracle["R_a_up"])
    # first init_N_len frames are (almost) silence: G = 1e-9 (bnmf_sep_event_RT_IS16.m:256-259)
    assert np.all(np.abs(out[:160 * 10].astype(int)) <= 1)


@pytest.mark.parametrize("wrong", [dict(alpha_d=0.7), dict(adapt_train_N=0), dict(beta=0.5), dict(alpha_p=0.5)],
                         ids=lambda d: "-".join(f"{k}={v}" for k, v in d.items()))
def test_golden_wav_pin_detects_wrong_parameters(bases, wavs, rng_inputs, wrong):
    """Negative controls: what the end-to-end pin CAN see.  A wrong noise-smoothing constant, no adaptation, a wrong
    over-subtraction weight or a wrong a-priori SNR smoothing all fall below the 21 dB the shipped settings reach.
    (It cannot see max_iter, conv_eps, sparsity, alpha_eta or blk_gap: their effect is inside the 21.7 .. 24.1 dB spread
    of the RNG stream.  Parity with the reference at kernel granularity needs tests/golden/make_ref_vectors.m.)"""
    p = O.default_params()
    p.update(wrong)
    h_init, Ad = rng_inputs
    out, _ = O.enhance_utterance(wavs["M03_in"], p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad)
    assert snr_db(wavs["M03_ref_out"], out) < 21.0


def test_oracle_matches_reference_golden_wav_LM_in(bases, wavs, rng_inputs):
    """The second wav pair the reference ships (LM_in, 17.7 s): length rule and the coarse SNR pin."""
    p = O.default_params()
    h_init, Ad = rng_inputs
    out, _ = O.enhance_utterance(wavs["LM_in"], p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad)
    assert len(out) == len(wavs["LM_ref_out"])
    assert snr_db(wavs["LM_ref_out"], out) > 16.0


def test_oracle_matches_reference_vectors(bases, wavs):
    """Hop-level pin against the UNMODIFIED reference run under Octave / MATLAB with the RNG replaced by the committed
    stream (tests/golden/make_ref_vectors.m + ref_shadow/rand.m).  Skipped until somebody with an interpreter commits
    tests/golden/ref_vectors.mat: neither exists in the build image nor on the GPU boxes (profiles/r02_octave_probe.txt)."""
    from conftest import GOLDEN, rel_err
    f = GOLDEN / "ref_vectors.mat"
    if not f.exists():
        pytest.skip("tests/golden/ref_vectors.mat not generated (needs GNU Octave or MATLAB)")
    import scipy.io
    ref = scipy.io.loadmat(f, squeeze_me=True, struct_as_record=False)["ref"]
    p = O.default_params()
    h_init, Ad, A_d = O.default_rng_inputs(p, with_A_d=True)
    tr = []
    out, g = O.enhance_utterance(wavs["M03_in"], p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad,
                                 A_d_init=A_d, trace=tr)
    assert len(tr) == int(ref.hops)
    Xt = np.stack([t["Xm_tilde"] for t in tr], axis=1)
    Ad_hop = np.stack([t["A"][p["R_x"]:] for t in tr], axis=1)
    assert rel_err(ref.Xm_tilde, Xt) <= 1e-3          # north_star: spectra within 1e-3
    assert rel_err(ref.A_d, Ad_hop) <= 1e-3           # activations within 1e-3
    assert rel_err(ref.B_DFT_d_final, g.B_DFT_d) <= 1e-3
    assert snr_db(np.asarray(ref.out_pcm, dtype=np.float64), out) >= 40.0   # waveform >= 40 dB


def test_oracle_lm_in_length_rule(wavs):
    p = O.default_params()
    assert O.num_hops(len(wavs["LM_in"]), p) - p["delay"] == len(wavs["LM_ref_out"]) // 160 == 1774


@pytest.mark.parametrize("n", [0, 1, 159, 160, 161, 1000])
def test_output_length_rule(bases, rng_inputs, n):
    p = O.default_params()
    h_init, Ad = rng_inputs
    pcm = (np.arange(n) % 200 - 100).astype(np.int16)
    out, _ = O.enhance_utterance(pcm, p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad)
    assert len(out) == (n // 160 + 1) * 160


def test_sparse_nmf_h_update_monotone_and_early_stop(bases, rng_inputs):
    rs = np.random.RandomState(0)
    W = np.concatenate([bases["B_DFT_x"], bases["B_DFT_d"]], axis=1)
    v = W @ rs.gamma(0.5, 1.0, size=(200, 3)) * 1e6 + 1e-9
    w, h, obj = O.sparse_nmf(v, init_w=W, init_h=rs.rand(200, 3), w_update_ind=np.zeros(200, bool),
                             h_update_ind=np.ones(200, bool), sparsity=5.0, conv_eps=1e-3, max_iter=100)
    cost = obj["cost"]
    assert len(cost) == obj["iters"] <= 100
    assert np.all(np.diff(cost) <= 1e-9 * np.abs(cost[:-1]))       # KL MU with fixed W never increases the cost
    assert abs(cost[-1] - cost[-2]) / cost[-2] < 1e-3               # stopped by the relative-change test
    assert np.allclose(np.linalg.norm(w, axis=0), 1.0, atol=1e-12)  # W returned normalised (:157-160)


def test_sparse_nmf_w_update_unit_columns(bases):
    rs = np.random.RandomState(1)
    V = rs.gamma(1.0, 1.0, size=(513, 100)) + 1e-9
    w, h, obj = O.sparse_nmf(V, init_w=bases["B_DFT_d"][:, :30], init_h=rs.rand(30, 100),
                             w_update_ind=np.ones(30, bool), h_update_ind=np.zeros(30, bool), sparsity=5.0,
                             conv_eps=1e-3, max_iter=50)
    assert np.allclose(np.linalg.norm(w, axis=0), 1.0, atol=1e-12)
    assert np.all(w >= 0)


def test_sparse_nmf_other_divergences_run():
    rs = np.random.RandomState(2)
    V = rs.rand(40, 30) + 0.1
    for cf in ("is", "ed"):
        w, h, obj = O.sparse_nmf(V, init_w=rs.rand(40, 5), init_h=rs.rand(5, 30), cf=cf, sparsity=0.1, max_iter=20)
        assert np.all(np.isfinite(w)) and np.all(np.isfinite(h))
        assert obj["div"][-1] < obj["div"][0]


def test_blk_sparse_properties():
    p = O.default_params()
    rs = np.random.RandomState(3)
    r_blk = rs.rand(513, 20)
    X, D = rs.rand(513) + 0.1, rs.rand(513) + 0.1
    Q, r_out = O.blk_sparse(X, D, r_blk, 25, p)
    assert Q.shape == (513,) and r_out.shape == (513, 20)
    assert np.all(Q >= 0) and np.all(Q <= 1)
    assert np.all(Q[:5] == 0)
    assert np.all(Q[5:59] == Q[64])                 # Q(1:59) = Q(65)   blk_sparse.m:32
    assert np.all(Q[483:] == 0.1)                   # untouched tail keeps the initial 0.1
    assert np.array_equal(r_out[:, :-1], r_blk[:, 1:])
    assert np.isclose(r_out[:, -1].max(), 1.0)
    Q0, _ = O.blk_sparse(X, D, r_blk, 20, p)        # before frame 21: default map
    assert np.all(Q0[5:] == 0.1) and np.all(Q0[:5] == 0)


def test_stft_istft_roundtrip():
    p = O.default_params()
    rs = np.random.RandomState(4)
    y = rs.randn(640) * 1000
    Ym, Yp = O.frame_stft(y, p)
    Ym0 = Ym.copy()
    s = O.synth_ifft_buff(Ym, Yp, 640, 1024, np.ones(640), 0.0, 0, 2.0)[:, 0]
    # with DCbin_back=0 only the analysis DC-zeroing (bins 1..5) is lost
    Y = np.fft.rfft(np.concatenate([p["win_STFT"] * y, np.zeros(384)]))
    Y[:5] = 0
    ref = np.fft.irfft(Y)[:640]
    assert np.allclose(s, ref, atol=1e-6 * np.abs(ref).max())
    assert np.all(Ym0[:5] == 1e-9)


def test_training_stft_quirks():
    p = O.default_params()
    s = np.random.RandomState(5).randn(16000) * 1000
    mag, ph = O.stft_fft(s, 640, 160, 1024, 5, p["win_STFT"], 0.0)
    assert mag.shape == (513, 100)
    n_valid = int(np.sum(np.any(mag != 0, axis=0)))
    assert n_valid == len(range(1, 16000 - 1024, 160))     # loop bound drops the tail (stft_fft.m:21)
    assert np.all(mag[:5, :n_valid] == 1e-6)                  # DC bins := 1e-6 (:31)


def test_mel_matrix_shape_and_coverage():
    M = O.mel_matrix(16000, 64, 1024, 1, 8000)
    assert M.shape == (513, 64)
    assert np.all(M >= 0) and np.all(M <= 1)
    assert np.all(M.sum(axis=0) > 0)


def test_snmf_mdi_keeps_observed_part():
    rs = np.random.RandomState(6)
    W = rs.rand(50, 8)
    V = W @ rs.rand(8, 12) + 1e-9
    Dm = (rs.rand(50, 12) > 0.3).astype(float)
    v_mdi, h, obj = O.snmf_mdi(V, Dm, init_w=W, init_h=rs.rand(8, 12), w_update_ind=np.zeros(8, bool),
                               h_update_ind=np.ones(8, bool), sparsity_mdi=0.0, conv_eps_mdi=1e-6, max_iter=200)
    assert np.allclose(v_mdi[Dm == 1], np.maximum(V, 1e-9)[Dm == 1])
    assert O_rel(V[Dm == 0], v_mdi[Dm == 0]) < 0.05


def O_rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(a)


def test_semi_supervised_switches_follow_the_reference_branch_order(bases, wavs, rng_inputs):
    """bnmf_sep_event_RT_IS16.m:125-139 is an if / elseif chain: with basis_update_N set, basis_update_E is never looked at
    (the `N && E` branch is unreachable); either switch changes the activations of the hop, the dictionaries of g do not
    move (the W of the solve is discarded, :150-154), and the joint solve still lowers the KL cost monotonically."""
    h_init, Ad = rng_inputs
    Bx, Bd = bases["B_DFT_x"], bases["B_DFT_d"]
    pcm = wavs["M03_in"][8000:8000 + 160 * 22].astype(float)
    outs = {}
    for name, over in (("sup", {}), ("N", dict(basis_update_N=1)), ("E", dict(basis_update_E=1)),
                       ("NE", dict(basis_update_N=1, basis_update_E=1))):
        p = dict(O.default_params(), adapt_train_N=0, **over)
        g = O.init_buff(Bx, Bd, Bx, Bd, p, Ad_blk_init=Ad)
        y = np.zeros(640)
        xs = []
        for l in range(1, 23):   # past the init_N_len = 15 noise-only frames
            y = np.concatenate([y[160:], pcm[(l - 1) * 160:l * 160]])
            _, _, xt, g = O.bnmf_sep_event_RT_IS16(y, l, g, p, h_init=h_init)
            xs.append(xt.copy())
        outs[name] = np.concatenate(xs)
        assert np.array_equal(g.B_DFT_d, Bd) and np.array_equal(g.B_DFT_x, Bx)
    assert np.array_equal(outs["NE"], outs["N"])
    assert not np.array_equal(outs["N"], outs["sup"]) and not np.array_equal(outs["E"], outs["sup"])
    assert not np.array_equal(outs["N"], outs["E"])
    # the joint W/H iteration on one frame: KL + sparsity cost is non-increasing (multiplicative updates)
    v = np.abs(np.random.RandomState(3).randn(513, 1)) ** 2 * 1e6 + 1e-9
    W = np.concatenate([Bx, Bd], axis=1)
    wi = np.concatenate([np.zeros(100, bool), np.ones(100, bool)])
    _, _, obj = O.sparse_nmf(v, init_w=W, init_h=h_init, max_iter=40, conv_eps=0.0, sparsity=5.0, w_update_ind=wi,
                             h_update_ind=np.ones(200, bool))
    assert np.all(np.diff(obj["cost"]) <= 1e-9 * obj["cost"][:-1])
