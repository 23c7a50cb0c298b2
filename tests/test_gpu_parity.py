"""GPU parity tests: the CUDA path (through the C ABI) against the float64 oracle on identical inputs.

Tolerances are the ones BASELINE.json:north_star states: H/W and enhanced spectra relative error <= 1e-3,
waveform SNR >= 40 dB.  The float64 kernels normally agree to ~1e-10; iteration counts and gates must be equal.
"""
import numpy as np
import pytest

from conftest import rel_err, snr_db

pytestmark = pytest.mark.gpu

SPEC_TOL = 1e-3     # north_star: spectra / activations relative error
WAVE_SNR_DB = 40.0  # north_star: waveform SNR


@pytest.fixture(scope="module")
def api():
    from se_snmf_nat_b200 import api as a
    return a


@pytest.fixture(scope="module")
def O():
    from oracle import snmf_oracle
    return snmf_oracle


def run_gpu_traced(api, p, pcms, bases, h_init, Ad, groups=None, **kw):
    ctx = api.get_context(0)
    b = api.Batch(ctx, p, bases["B_DFT_x"], bases["B_DFT_d"], [len(x) for x in pcms], h_init, Ad, **kw)
    if groups:
        b.set_groups(groups)
    b.enable_trace(True)
    b.upload(pcms)
    b.run()
    outs = b.download()
    return b, outs


def test_m03_full_parity(api, O, bases, wavs, rng_inputs, m03_oracle):
    """config 1: filewise_run_IS16 on wav/M03_423C0213_STR.CH6.wav with the shipped dictionaries."""
    h_init, Ad = rng_inputs
    p = api.default_p()
    b, outs = run_gpu_traced(api, p, [wavs["M03_in"]], bases, h_init, Ad)
    out = outs[0]
    ref = m03_oracle
    assert len(out) == len(ref["out"]) == 55040
    hi = b.trace(0, "h_iters").astype(int)
    first_bad = np.flatnonzero(hi != ref["h_iters"])
    assert first_bad.size == 0, f"H-solve iteration counts diverge first at hop {first_bad[:1]}"
    assert np.array_equal(b.trace(0, "gated").astype(int), ref["gated"].astype(int))
    assert np.array_equal(b.trace(0, "R_a_up").astype(int), ref["R_a_up"])
    assert np.array_equal(b.trace(0, "w_iters").astype(int), ref["w_iters"])
    A = b.trace(0, "A")
    Xt = b.trace(0, "Xm_tilde")
    worstA = max(rel_err(ref["A"][i], A[i]) for i in range(len(hi)))
    worstX = max(rel_err(ref["Xm_tilde"][i], Xt[i]) for i in range(len(hi)))
    assert worstA <= SPEC_TOL, worstA
    assert worstX <= SPEC_TOL, worstX
    assert rel_err(ref["B_DFT_d_final"], b.noise_basis(0)) <= SPEC_TOL
    assert snr_db(ref["out"], out) >= WAVE_SNR_DB
    # against the reference's own shipped output: the coarse end-to-end pin
    assert snr_db(wavs["M03_ref_out"], out) > 20.0
    st = b.stats()
    assert st["hops"] == len(hi) and st["h_iters"] == int(ref["h_iters"].sum())
    assert st["w_iters"] == int(ref["w_iters"].sum())
    print(f"M03: worst rel err A {worstA:.2e}, Xm_tilde {worstX:.2e}, SNR vs oracle {snr_db(ref['out'], out):.1f} dB")
    b.close()


def test_gpu_matches_reference_vectors(api, bases, wavs):
    """The CUDA path against the UNMODIFIED reference (tests/golden/ref_vectors.mat, see make_ref_vectors.m); skipped
    until the file exists (no Octave / MATLAB in the build image or on the GPU boxes)."""
    from conftest import GOLDEN
    f = GOLDEN / "ref_vectors.mat"
    if not f.exists():
        pytest.skip("tests/golden/ref_vectors.mat not generated (needs GNU Octave or MATLAB)")
    import scipy.io
    from oracle import snmf_oracle as Or
    ref = scipy.io.loadmat(f, squeeze_me=True, struct_as_record=False)["ref"]
    h_init, Ad = Or.default_rng_inputs(Or.default_params())
    b, outs = run_gpu_traced(api, api.default_p(), [wavs["M03_in"]], bases, h_init, Ad)
    assert rel_err(np.asarray(ref.Xm_tilde).T, b.trace(0, "Xm_tilde")) <= SPEC_TOL
    assert rel_err(np.asarray(ref.A_d).T, b.trace(0, "A")[:, 100:]) <= SPEC_TOL
    assert rel_err(ref.B_DFT_d_final, b.noise_basis(0)) <= SPEC_TOL
    assert snr_db(np.asarray(ref.out_pcm, dtype=np.float64), outs[0]) >= WAVE_SNR_DB
    b.close()


def test_ragged_batch_equals_single_runs(api, O, bases, wavs, rng_inputs):
    """Utterances of different length (including empty and shorter than one hop) in one batch give exactly what
    each gives alone, and what the oracle gives."""
    h_init, Ad = rng_inputs
    p = api.default_p()
    po = O.default_params()
    rs = np.random.RandomState(7)
    m04 = wavs["M04_in"]
    pcms = [m04[:16000], np.zeros(0, np.int16), m04[5000:5100], m04[20000:20000 + 160 * 37 + 13],
            (rs.randn(8000) * 3000).astype(np.int16)]
    ads = np.stack([rs.rand(50, 100) for _ in pcms])
    outs = api.enhance_batch(pcms, p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=ads)
    for i, pcm in enumerate(pcms):
        ref, _ = O.enhance_utterance(pcm, po, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=ads[i])
        assert len(outs[i]) == len(ref) == (len(pcm) // 160 + 1) * 160
        assert np.abs(outs[i].astype(int) - ref.astype(int)).max() <= 1, i
        alone = api.enhance_batch([pcm], p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=ads[i])[0]
        assert np.array_equal(alone, outs[i]), "batch composition changed a result"


def test_digital_silence_and_constant_input(api, O, bases, rng_inputs):
    """All-zero PCM (every bin at p.nonzerofloor: the H-solve never meets its stop rule and runs max_iter), a constant
    (DC only, zeroed by DCbin), and silence followed by a tone: extreme arguments for the reciprocal / logarithm of the
    solves and for the gates.  Same iteration counts, gates and PCM as the oracle, in one batch (multi-stream kernel off and
    on: 3 streams -> per-stream kernel; repeated 3x -> 9 streams in one group -> multi-stream kernel)."""
    h_init, Ad = rng_inputs
    p, po = api.default_p(), O.default_params()
    tone = (3000 * np.sin(np.arange(3200) * 0.3)).astype(np.int16)
    base = [np.zeros(4000, np.int16), np.full(4000, 7, np.int16), np.concatenate([np.zeros(3200, np.int16), tone])]
    refs = []
    for pcm in base:
        tr = []
        ref, _ = O.enhance_utterance(pcm, po, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad, trace=tr)
        refs.append((ref, tr))
    for rep in (1, 3):
        pcms = base * rep
        b, outs = run_gpu_traced(api, p, pcms, bases, h_init, np.stack([Ad] * len(pcms)), groups=1)
        for i in range(len(pcms)):
            ref, tr = refs[i % 3]
            assert np.array_equal(b.trace(i, "h_iters").astype(int), np.array([t["h_iters"] for t in tr])), (rep, i)
            assert np.array_equal(b.trace(i, "w_iters").astype(int), np.array([t["w_iters"] for t in tr])), (rep, i)
            assert np.array_equal(b.trace(i, "gated").astype(int), np.array([int(t["gated"]) for t in tr])), (rep, i)
            assert np.isfinite(b.trace(i, "A")).all()
            assert np.abs(outs[i].astype(int) - ref.astype(int)).max() <= 1, (rep, i)
        b.close()


@pytest.mark.parametrize("variant", ["wiener", "no_adapt", "no_blk", "maxiter25_gap5", "preemph", "R_a20_ma40",
                                     "overlap0.1_ma40", "overlap0.5_ma40", "event3"])
def test_settings_variants(api, O, bases, wavs, rng_inputs, variant):
    """The knobs the reference's settings/bak_IS16_results variants change are runtime parameters."""
    h_init, Ad = rng_inputs
    over = {
        "wiener": dict(ENHANCE_METHOD="Wiener"),
        "no_adapt": dict(adapt_train_N=0),
        "no_blk": dict(blk_sparse=0),
        "maxiter25_gap5": dict(max_iter=25, blk_gap=5),
        "preemph": dict(preemph=0.92),
        "R_a20_ma40": dict(R_a=20, m_a=40),
        # settings/initial_setting_SNMF.m:57-58: an update only every floor(overlap_m_a * m_a) = 4 / 20 gated hops
        # (the update_switch counting of bnmf_sep_event_RT_IS16.m:293,343-345)
        "overlap0.1_ma40": dict(m_a=40, overlap_m_a=0.1),
        "overlap0.5_ma40": dict(m_a=40, overlap_m_a=0.5),
        # settings/bak_IS16_results/initial_setting_Proposed_Techwin_201603_RT.m:40-49: three event classes
        "event3": dict(EVENT_NUM=3, EVENT_RANK=[1, 21, 41]),
    }[variant]
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    if variant == "R_a20_ma40":
        Ad = np.random.RandomState(11).rand(20, 40)
    if variant.startswith("overlap"):
        Ad = np.random.RandomState(12).rand(50, 40)
    pcm = wavs["M03_in"][8000:8000 + 160 * 120]
    out = api.enhance_batch([pcm], p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad)[0]
    ref, _ = O.enhance_utterance(pcm, po, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad)
    assert len(out) == len(ref)
    assert snr_db(ref, out) >= WAVE_SNR_DB
    assert np.abs(out.astype(int) - ref.astype(int)).max() <= 1


@pytest.mark.parametrize("mode", ["semisupervised_N", "update_E_shipped"])
def test_semi_supervised_separation_solve(api, O, bases, wavs, mode):
    """p.basis_update_N / _E (bnmf_sep_event_RT_IS16.m:125-139): the per-hop sparse_nmf also updates half of the dictionary
    while it iterates and the caller keeps only A.  `semisupervised_N` is settings/bak_IS16_results/
    initial_setting_semisupervised.m (R_d = 50, no adaptation, no block sparsity, Wiener, max_iter = 25, pre-emphasis 0.92,
    DCfreq 160 Hz -> DCbin 10); `update_E_shipped` flips the other switch on top of the shipped settings, adaptation on."""
    if mode == "semisupervised_N":
        over = dict(R_d=50, R_a=50, adapt_train_N=0, init_N_len=10, m_a=40, overlap_m_a=0.5, blk_sparse=0, P_len_k=50,
                    P_len_l=3, alpha_p=0.6, preemph=0.92, DCbin=10, DCbin_back=10, max_iter=25, basis_update_N=1,
                    ENHANCE_METHOD="Wiener", alpha_eta=0.95, alpha_d=0.85, beta=2.0)
        Bd = bases["B_DFT_d"][:, :50].copy()
        hops = 60
    else:
        over = dict(basis_update_E=1)
        Bd = bases["B_DFT_d"]
        hops = 40
    Bx = bases["B_DFT_x"]
    R = Bx.shape[1] + Bd.shape[1]
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    h_init = O.park_miller(R, 1)
    Ad = np.random.RandomState(5).rand(p["R_a"], p["m_a"])
    pcms = [wavs["M03_in"][8000:8000 + 160 * hops], wavs["M04_in"][3000:3000 + 160 * (hops - 7) + 55]]
    ctx = api.get_context(0)
    b = api.Batch(ctx, p, Bx, Bd, [len(x) for x in pcms], h_init, np.stack([Ad, Ad]))
    b.enable_trace(True)
    b.upload(pcms)
    b.run()
    outs = b.download()
    for i, pcm in enumerate(pcms):
        tr = []
        ref, _ = O.enhance_utterance(pcm, po, Bx, Bd, h_init=h_init, Ad_blk_init=Ad, trace=tr)
        assert np.array_equal(b.trace(i, "h_iters").astype(int), np.array([t["h_iters"] for t in tr])), i
        assert np.array_equal(b.trace(i, "w_iters").astype(int), np.array([t["w_iters"] for t in tr])), i
        A = b.trace(i, "A")
        assert max(rel_err(tr[k]["A"], A[k]) for k in range(len(tr))) <= SPEC_TOL
        assert snr_db(ref, outs[i]) >= WAVE_SNR_DB
        assert np.abs(outs[i].astype(int) - ref.astype(int)).max() <= 1, i
    b.close()
    # the solve really differs from the supervised one: same inputs, switch off -> other activations
    p0 = dict(p, basis_update_N=0, basis_update_E=0)
    sup = api.enhance_batch([pcms[0]], p0, Bx, Bd, h_init=h_init, Ad_blk_init=Ad)[0]
    assert not np.array_equal(sup, outs[0])


def _grow(B, cols, seed):
    """A dictionary with `cols` columns made of perturbed copies of the shipped one (unit-free, non-negative)."""
    rs = np.random.RandomState(seed)
    reps = [B[:, rs.permutation(B.shape[1])] * np.exp(0.3 * rs.randn(*B.shape)) for _ in range((cols + B.shape[1] - 1) // B.shape[1])]
    return np.concatenate([B] + reps, axis=1)[:, :cols].copy()


@pytest.mark.parametrize("setting", ["techwin_R240_3classes", "exemplar_R1000"])
def test_large_rank_settings_of_the_reference(api, O, bases, wavs, rng_inputs, setting):
    """settings/bak_IS16_results/initial_setting_Proposed_Techwin_201603_RT.m:40-61 (R_x = 140 in three event classes,
    R_d = 100, R_a = 25, overlap_m_a = 0.1, max_iter = 25) and initial_setting_Exemplar.m:47-58,104 (R_x = R_d = 500, no
    adaptation, max_iter = 50): dictionaries that do not fit the cluster-resident H-solve run on the streaming kernel and
    must still reproduce the oracle hop by hop."""
    if setting.startswith("techwin"):
        over = dict(EVENT_NUM=3, EVENT_RANK=[1, 21, 41], R_x=140, R_d=100, R_a=25, m_a=100, overlap_m_a=0.1, max_iter=25)
        Bx, Bd = _grow(bases["B_DFT_x"], 140, 1), bases["B_DFT_d"]
        hops = 70
    else:
        over = dict(R_x=500, R_d=500, R_a=50, m_a=40, overlap_m_a=0.5, adapt_train_N=0, max_iter=50)
        Bx, Bd = _grow(bases["B_DFT_x"], 500, 2), _grow(bases["B_DFT_d"], 500, 3)
        hops = 25
    R = over["R_x"] + over["R_d"]
    rs = np.random.RandomState(41)
    h_init = O.park_miller(R, 1)
    Ad = rs.rand(over["R_a"], over["m_a"])
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    pcms = [wavs["M03_in"][8000:8000 + 160 * hops], wavs["M04_in"][3000:3000 + 160 * (hops - 9) + 31]]
    ctx = api.get_context(0)
    b = api.Batch(ctx, p, Bx, Bd, [len(x) for x in pcms], h_init, np.stack([Ad, Ad]))
    b.enable_trace(True)
    b.upload(pcms)
    b.run()
    outs = b.download()
    for i, pcm in enumerate(pcms):
        tr = []
        ref, _ = O.enhance_utterance(pcm, po, Bx, Bd, h_init=h_init, Ad_blk_init=Ad, trace=tr)
        assert np.array_equal(b.trace(i, "h_iters").astype(int), np.array([t["h_iters"] for t in tr])), i
        assert np.array_equal(b.trace(i, "w_iters").astype(int), np.array([t["w_iters"] for t in tr])), i
        A = b.trace(i, "A")
        assert max(rel_err(tr[k]["A"], A[k]) for k in range(len(tr))) <= SPEC_TOL
        assert np.abs(outs[i].astype(int) - ref.astype(int)).max() <= 1, i
    b.close()


def test_chain_mode_carries_the_noise_basis_between_files(api, O, bases, wavs, rng_inputs):
    """Do_MultiBatch / NTF_sep_event_RT semantics (src/NTF_sep_event_RT.m:28-38,136-139): inside a target directory the
    adapted noise basis of file i is the starting basis of file i+1 (B_D_u.mat); chains are independent of each other
    and of the unchained utterances in the same batch."""
    h_init, _ = rng_inputs
    p = api.default_p()
    po = O.default_params()
    rs = np.random.RandomState(21)
    m03, m04 = wavs["M03_in"], wavs["M04_in"]
    pcms = [m03[2000:2000 + 160 * 70], m04[:160 * 45], m03[20000:20000 + 160 * 55 + 40], m04[30000:30000 + 160 * 80],
            m03[40000:40000 + 160 * 30], np.zeros(0, np.int16), m04[50000:50000 + 160 * 25]]
    chain = [0, 7, 0, 7, -1, 0, 7]          # chain 0: files 0,2,5 ; chain 7: files 1,3,6 ; file 4 alone
    ads = np.stack([rs.rand(50, 100) for _ in pcms])
    ctx = api.get_context(0)
    b = api.Batch(ctx, p, bases["B_DFT_x"], bases["B_DFT_d"], [len(x) for x in pcms], h_init, ads, chain_id=chain)
    b.upload(pcms)
    b.run()
    outs = b.download()
    for members in ([0, 2, 5], [1, 3, 6], [4]):
        ref, Bd = O.enhance_chain([pcms[i] for i in members], po, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init,
                                  Ad_blk_inits=[ads[i] for i in members])
        for i, r in zip(members, ref):
            assert len(outs[i]) == len(r)
            assert np.abs(outs[i].astype(int) - r.astype(int)).max() <= 1, i
        assert rel_err(Bd, b.noise_basis(members[-1])) <= SPEC_TOL      # what the reference would leave in B_D_u.mat
    # a second run of the same batch object starts every chain from the shipped basis again
    b.run()
    again = b.download()
    assert all(np.array_equal(a, o) for a, o in zip(again, outs))
    # chaining matters: file 2 enhanced on its own differs from file 2 as the second file of chain 0
    alone = api.enhance_batch([pcms[2]], p, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=ads[2])[0]
    assert not np.array_equal(alone, outs[2])
    b.close()


def test_multi_stream_hsolve_matches_oracle(api, O, bases, wavs, rng_inputs):
    """17 ragged utterances in ONE slot group: the multi-stream H-solve (7 streams in lock step per 8-CTA cluster, the 150
    stream-invariant columns as FP64 tensor-core fragments; clusters of 7, 7 and 3 streams, sorted launch order) runs
    while >= 7 streams are active, the per-stream kernel afterwards.  Every utterance must match the oracle hop by hop:
    iteration counts, gates, activations, waveform."""
    h_init, _ = rng_inputs
    p = api.default_p()
    po = O.default_params()
    rs = np.random.RandomState(11)
    src = np.concatenate([wavs["M04_in"], wavs["M03_in"]])
    pcms = []
    for i in range(17):
        n = 160 * int(rs.randint(22, 60)) + int(rs.randint(0, 160))
        o = int(rs.randint(0, len(src) - n))
        pcms.append(src[o:o + n])
    ads = np.stack([rs.rand(50, 100) for _ in pcms])
    b, outs = run_gpu_traced(api, p, pcms, bases, h_init, ads, groups=1)
    assert b.stats()["ms_launches"] > 0 if "ms_launches" in b.stats() else True
    tot_h = 0
    for i, pcm in enumerate(pcms):
        tr = []
        ref, _ = O.enhance_utterance(pcm, po, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=ads[i], trace=tr)
        hi = b.trace(i, "h_iters").astype(int)
        assert np.array_equal(hi, np.array([t["h_iters"] for t in tr])), i
        assert np.array_equal(b.trace(i, "gated").astype(int), np.array([int(t["gated"]) for t in tr])), i
        assert np.array_equal(b.trace(i, "w_iters").astype(int), np.array([t["w_iters"] for t in tr])), i
        A = b.trace(i, "A")
        assert max(rel_err(tr[k]["A"], A[k]) for k in range(len(tr))) <= SPEC_TOL, i
        assert np.abs(outs[i].astype(int) - ref.astype(int)).max() <= 1, i
        tot_h += int(hi.sum())
    assert b.stats()["h_iters"] == tot_h
    b.close()


def test_stream_groups_do_not_change_results(api, bases, wavs, rng_inputs):
    """snmfnat_batch_set_groups is a scheduling knob: interleaved slot groups on separate CUDA streams give bit-identical
    output (also with chains, whose boundary re-initialisation has to run on the owning group's stream)."""
    h_init, _ = rng_inputs
    p = api.default_p()
    rs = np.random.RandomState(3)
    m04 = wavs["M04_in"]
    pcms = [m04[o:o + n] for o, n in [(0, 160 * 40), (7000, 160 * 33 + 5), (15000, 160 * 52), (26000, 160 * 18),
                                       (31000, 160 * 47), (40000, 160 * 29), (47000, 160 * 36)]]
    chain = [0, -1, 0, 1, -1, 1, 0]
    ads = np.stack([rs.rand(50, 100) for _ in pcms])
    ctx = api.get_context(0)
    res = {}
    for ng in (1, 3):
        b = api.Batch(ctx, p, bases["B_DFT_x"], bases["B_DFT_d"], [len(x) for x in pcms], h_init, ads, chain_id=chain)
        b.set_groups(ng)
        b.upload(pcms)
        b.run()
        res[ng] = (b.download(), b.stats())
        b.close()
    assert all(np.array_equal(a, c) for a, c in zip(res[1][0], res[3][0]))
    for k in ("hops", "h_iters", "w_iters", "w_solves"):
        assert res[1][1][k] == res[3][1][k]


def test_mel_separation_mode(api, O, bases, wavs, rng_inputs):
    """p.B_sep_mode = 'Mel' (SURVEY.md 8f rank 1): separation and adaptation on the 64-band Mel dictionaries
    (B_Mel_sub of the shipped basis files), gain / block sparsity / resynthesis in the DFT domain
    (bnmf_sep_event_RT_IS16.m:107-119,165-211,295-319; init_buff.m:60-62)."""
    h_init, Ad = rng_inputs
    over = dict(B_sep_mode="Mel", MelConv=1)
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    pcms = [wavs["M03_in"][6000:6000 + 160 * 130], wavs["M04_in"][12000:12000 + 160 * 60 + 9]]
    ads = np.stack([Ad, np.random.RandomState(4).rand(50, 100)])
    ctx = api.get_context(0)
    b = api.Batch(ctx, p, bases["B_DFT_x"], bases["B_DFT_d"], [len(x) for x in pcms], h_init, ads,
                  B_Mel_x=bases["B_Mel_x"], B_Mel_d=bases["B_Mel_d"])
    b.enable_trace(True)
    b.upload(pcms)
    b.run()
    outs = b.download()
    for i, pcm in enumerate(pcms):
        tr = []
        ref, g = O.enhance_utterance(pcm, po, bases["B_DFT_x"], bases["B_DFT_d"], h_init=h_init, Ad_blk_init=ads[i],
                                     B_Mel_x=bases["B_Mel_x"], B_Mel_d=bases["B_Mel_d"], trace=tr)
        assert len(outs[i]) == len(ref)
        assert np.array_equal(b.trace(i, "h_iters").astype(int), np.array([t["h_iters"] for t in tr]))
        assert np.array_equal(b.trace(i, "w_iters").astype(int), np.array([t["w_iters"] for t in tr]))
        assert sum(t["w_iters"] for t in tr) > 0                       # the Mel-domain adaptation did run
        worst = max(rel_err(t["Xm_tilde"], x) for t, x in zip(tr, b.trace(i, "Xm_tilde")))
        assert worst <= SPEC_TOL, worst
        assert snr_db(ref, outs[i]) >= WAVE_SNR_DB
        assert np.abs(outs[i].astype(int) - ref.astype(int)).max() <= 1
        Bm = b.noise_basis(i)
        assert Bm.shape == (64, 100)
        assert rel_err(g.B_Mel_d, Bm) <= SPEC_TOL
    b.close()
    # Mel mode without the Mel dictionaries, or with MelConv = 0, is refused
    with pytest.raises(ValueError):
        api.Batch(ctx, p, bases["B_DFT_x"], bases["B_DFT_d"], [1000], h_init, Ad)
    with pytest.raises(api.SnmfnatError):
        api.Batch(ctx, dict(p, MelConv=0), bases["B_DFT_x"], bases["B_DFT_d"], [1000], h_init, Ad,
                  B_Mel_x=bases["B_Mel_x"], B_Mel_d=bases["B_Mel_d"])


def test_unsupported_configs_fail_loudly(api, bases, rng_inputs):
    h_init, Ad = rng_inputs
    for over in (dict(Splice=1), dict(blk_len_sep=2, blk_hop_sep=2)):
        with pytest.raises(api.SnmfnatError) as e:
            api.enhance_batch([np.zeros(1000, np.int16)], dict(api.default_p(), **over), bases["B_DFT_x"],
                              bases["B_DFT_d"], h_init=h_init, Ad_blk_init=Ad)
        assert e.value.code == -4


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


_VARIANT_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
from se_snmf_nat_b200 import api
from oracle import snmf_oracle as O
g = {golden!r}
bases = np.load(g + "/bases.npz"); wavs = np.load(g + "/wavs.npz"); rng = np.load(g + "/rng_seed1.npz")
n = int(os.environ.get("SNMFNAT_TEST_NUTT", "1"))
hops = 150 if n == 1 else 60
pcms = [wavs["M03_in"][4000 + 3000 * i:4000 + 3000 * i + 160 * (hops - 3 * i)] for i in range(n)]
ads = np.stack([rng["Ad_blk"]] * n)
outs, st = api.enhance_batch(pcms, api.default_p(), bases["B_DFT_x"], bases["B_DFT_d"], h_init=rng["h_init"],
                             Ad_blk_init=ads, return_stats=True)
hi = wi = d = 0
for pcm, out in zip(pcms, outs):
    tr = []
    ref, _ = O.enhance_utterance(pcm, O.default_params(), bases["B_DFT_x"], bases["B_DFT_d"], h_init=rng["h_init"],
                                 Ad_blk_init=rng["Ad_blk"], trace=tr)
    assert len(out) == len(ref)
    d = max(d, int(np.abs(out.astype(int) - ref.astype(int)).max()))
    hi += sum(int(t["h_iters"]) for t in tr); wi += sum(int(t["w_iters"]) for t in tr)
assert d <= 1, d
assert st["h_iters"] == hi, (st["h_iters"], hi)
assert st["w_iters"] == wi, (st["w_iters"], wi)
print("variant ok", d, st["h_iters"], st["w_iters"])
"""


@pytest.mark.parametrize("env", [dict(SNMFNAT_FORCE_GENERIC="1"), dict(SNMFNAT_HSOLVE="ms"), dict(SNMFNAT_HSOLVE="single"),
                                 dict(SNMFNAT_HSOLVE="ms7", SNMFNAT_GROUPS="1", SNMFNAT_TEST_NUTT="9")],
                         ids=["generic_kernels", "hsolve_ms_forced", "hsolve_single_forced", "hsolve_ms7_shared_memory_only"])
def test_alternative_kernel_generations_agree_with_oracle(env):
    """SNMFNAT_FORCE_GENERIC=1 selects the any-geometry kernels (the ones every non-shipped rank/frame-length falls back
    to); SNMFNAT_HSOLVE=ms / single force the multi-stream (one live stream in a 7-stream cluster) / the per-stream H-solve
    whatever the number of active streams; SNMFNAT_HSOLVE=ms7 selects the 7-stream all-shared-memory variant of the
    multi-stream kernel instead of the 8-stream one with four streams in tensor memory (it needs >= 7 active streams, so that
    case runs a 9-utterance batch).  All must reproduce the oracle.  The switches are read once per process."""
    import os
    import subprocess
    import sys
    from conftest import GOLDEN, ROOT
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT.format(root=str(ROOT), golden=str(GOLDEN))], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "variant ok" in r.stdout
