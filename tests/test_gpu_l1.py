"""GPU parity of the L1 / L2 entry points (sparse_nmf, snmf_mdi, DNMF_adapt, stft_fft, synth_ifft_buff, blk_sparse,
init_buff + bnmf_sep_event_RT_IS16) against the float64 oracle, through the C ABI."""
import numpy as np
import pytest

from conftest import rel_err, snr_db

pytestmark = pytest.mark.gpu
TOL = 1e-3   # north_star: H/W relative error; float64 kernels are expected near 1e-12


@pytest.fixture(scope="module")
def api():
    from se_snmf_nat_b200 import api as a
    return a


@pytest.fixture(scope="module")
def O():
    from oracle import snmf_oracle
    return snmf_oracle


def _rand_from(rs):
    return lambda m, n: rs.rand(n, m).T.copy()     # column-major fill like MATLAB's rand(m, n)


@pytest.mark.parametrize("cf,wu,hu", [("kl", True, True), ("kl", False, True), ("kl", True, False), ("ed", True, True),
                                      ("is", True, True), ("beta1.5", True, True)])
def test_sparse_nmf_matches_oracle(api, O, cf, wu, hu):
    rs = np.random.RandomState(0)
    F, n, r = 97, 45, 12
    V = rs.gamma(1.0, 1.0, size=(F, n)) + 1e-3
    w0, h0 = rs.rand(F, r) + 0.05, rs.rand(r, n) + 0.05
    p = dict(init_w=w0, init_h=h0, max_iter=40, conv_eps=1e-4, sparsity=0.3, cost_check=1, cf=cf if cf != "beta1.5" else "x",
             w_update_ind=np.full(r, wu), h_update_ind=np.full(r, hu))
    kw = dict(cf=cf) if cf != "beta1.5" else dict(cf="x", beta=1.5)
    if cf == "beta1.5":
        p["beta"] = 1.5
    w, h, obj = api.sparse_nmf(V, p)
    wo, ho, oo = O.sparse_nmf(V, init_w=w0, init_h=h0, max_iter=40, conv_eps=1e-4, sparsity=0.3,
                              w_update_ind=np.full(r, wu), h_update_ind=np.full(r, hu), **kw)
    assert obj["iters"] == oo["iters"]
    assert rel_err(wo, w) < 1e-9 and rel_err(ho, h) < 1e-9
    assert np.allclose(obj["cost"], oo["cost"], rtol=1e-10)
    assert np.allclose(obj["div"], oo["div"], rtol=1e-10)


def test_sparse_nmf_partial_indices_and_matrix_sparsity(api, O):
    rs = np.random.RandomState(1)
    F, n, r = 64, 30, 10
    V = rs.gamma(1.0, 1.0, size=(F, n)) + 1e-3
    w0, h0 = rs.rand(F, r) + 0.05, rs.rand(r, n) + 0.05
    wi = np.arange(r) % 2 == 0
    hi = np.arange(r) < 7
    sp = rs.rand(r, n)
    p = dict(init_w=w0, init_h=h0, max_iter=25, conv_eps=0, sparsity=sp, cost_check=1, w_update_ind=wi, h_update_ind=hi)
    w, h, obj = api.sparse_nmf(V, p)
    wo, ho, oo = O.sparse_nmf(V, init_w=w0, init_h=h0, max_iter=25, conv_eps=0, sparsity=sp, w_update_ind=wi, h_update_ind=hi)
    assert obj["iters"] == 25 and rel_err(wo, w) < 1e-9 and rel_err(ho, h) < 1e-9
    # rows of h outside h_ind only get the initial rescaling by the column norms
    assert np.allclose(h[7:], h0[7:] * np.linalg.norm(w0, axis=0)[7:, None])


def test_sparse_nmf_requires_inits_and_cost_check(api):
    V = np.ones((8, 4))
    with pytest.raises(KeyError):
        api.sparse_nmf(V, dict(init_w=np.ones((8, 2)), init_h=np.ones((2, 4))))
    with pytest.raises(ValueError):
        api.sparse_nmf(V, dict(cost_check=1))
    with pytest.raises(ValueError):
        api.sparse_nmf(V, dict(cost_check=1, r=3))      # no rand= given


def test_sparse_nmf_online_shapes(api, O, bases, rng_inputs):
    """The two shapes of the online path through the generic entry: 513 x 1 H-solve and 513 x 100 W-solve."""
    h_init, Ad = rng_inputs
    rs = np.random.RandomState(2)
    W = np.concatenate([bases["B_DFT_x"], bases["B_DFT_d"]], axis=1)
    v = W @ rs.gamma(0.5, 1.0, size=(200, 1)) * 1e6 + 1e-9
    p = dict(init_w=W, init_h=h_init, max_iter=100, conv_eps=1e-3, sparsity=5.0, cost_check=1,
             w_update_ind=np.zeros(200, bool), h_update_ind=np.ones(200, bool))
    w, h, obj = api.sparse_nmf(v, p)
    wo, ho, oo = O.sparse_nmf(v, init_w=W, init_h=h_init, max_iter=100, conv_eps=1e-3, sparsity=5.0,
                              w_update_ind=np.zeros(200, bool), h_update_ind=np.ones(200, bool))
    assert obj["iters"] == oo["iters"] and rel_err(ho, h) < 1e-9
    V = rs.gamma(1.0, 1e5, size=(513, 100))
    p = dict(init_w=bases["B_DFT_d"][:, :41], init_h=Ad[:41], max_iter=100, conv_eps=1e-3, sparsity=5.0, cost_check=1,
             w_update_ind=np.ones(41, bool), h_update_ind=np.zeros(41, bool))
    w, h, obj = api.sparse_nmf(V, p)
    wo, ho, oo = O.sparse_nmf(V, init_w=bases["B_DFT_d"][:, :41], init_h=Ad[:41], max_iter=100, conv_eps=1e-3,
                              sparsity=5.0, w_update_ind=np.ones(41, bool), h_update_ind=np.zeros(41, bool))
    assert obj["iters"] == oo["iters"] and rel_err(wo, w) < 1e-9
    assert np.allclose(np.linalg.norm(w, axis=0), 1.0, atol=1e-12)


@pytest.mark.parametrize("soft", [False, True])
def test_snmf_mdi(api, O, soft):
    rs = np.random.RandomState(3)
    W = rs.rand(50, 8)
    V = W @ rs.rand(8, 12) + 1e-9
    Dm = rs.rand(50, 12) if soft else (rs.rand(50, 12) > 0.3).astype(float)
    h0 = rs.rand(8, 12)
    p = dict(init_w=W, init_h=h0, w_update_ind=np.zeros(8, bool), h_update_ind=np.ones(8, bool), sparsity_mdi=0.1,
             conv_eps_mdi=1e-6, max_iter=60, cost_check=1)
    vm, h, obj = api.snmf_mdi(V, Dm, p, soft=soft)
    vo, ho, oo = O.snmf_mdi(V, Dm, soft_mask=soft, init_w=W, init_h=h0, w_update_ind=np.zeros(8, bool),
                            h_update_ind=np.ones(8, bool), sparsity_mdi=0.1, conv_eps_mdi=1e-6, max_iter=60)
    assert obj["iters"] == oo["iters"]
    assert rel_err(vo, vm) < 1e-9 and rel_err(ho, h) < 1e-9
    with pytest.raises(KeyError):
        api.snmf_mdi(V, Dm, dict(init_w=W, init_h=h0, cost_check=1))


def test_dnmf_adapt(api, O):
    rs = np.random.RandomState(4)
    F, n, Rx, Rd = 80, 40, 6, 5
    B = rs.rand(F, Rx + Rd) + 0.01
    Y = B @ rs.gamma(0.5, 1.0, size=(Rx + Rd, n)) + 1e-6
    D = B[:, Rx:] @ rs.gamma(0.5, 1.0, size=(Rd, n)) + 1e-6
    p = dict(R_x=Rx, R_d=Rd, max_iter=30, conv_eps=1e-4, sparsity=0.5, cost_check=1)
    Ba = api.DNMF_adapt(Y, D, B, p, rand=_rand_from(np.random.RandomState(9)))
    Bo = O.dnmf_adapt(Y, D, B, R_x=Rx, R_d=Rd, rand=_rand_from(np.random.RandomState(9)), max_iter=30, conv_eps=1e-4,
                      sparsity=0.5)
    assert Ba.shape == (F, Rd) and rel_err(Bo, Ba) < 1e-9


@pytest.mark.parametrize("preemph", [0.0, 0.92])
def test_stft_fft_and_synth(api, O, preemph):
    po = O.default_params()
    s = np.random.RandomState(5).randn(16000 * 2 + 77) * 1000
    mag, ph = api.stft_fft(s, 640, 160, 1024, 5, po["win_STFT"], preemph)
    mo, pho = O.stft_fft(s, 640, 160, 1024, 5, po["win_STFT"], preemph)
    assert mag.shape == mo.shape
    assert rel_err(mo, mag) < 1e-12
    nv = int(np.sum(np.any(mo != 0, axis=0)))
    assert np.all(mag[:, nv:] == 0) and np.all(ph[:, nv:] == 0)      # trailing preallocated columns stay zero
    d = np.angle(np.exp(1j * (ph[:, :nv] - pho[:, :nv])))
    assert np.max(np.abs(d[mo[:, :nv] > 1e-3])) < 1e-9
    # resynthesis of 50 frames, half-spectrum and full-spectrum input forms (synth_ifft_buff.m:13-19)
    P = mo[:, :50] ** 2
    sb = api.synth_ifft_buff(P, pho[:, :50], 640, 1024, po["win_ISTFT"], preemph, 5, 2.0)
    so = O.synth_ifft_buff(P, pho[:, :50], 640, 1024, po["win_ISTFT"], preemph, 5, 2.0)
    assert rel_err(so, sb) < 1e-10
    full = np.random.RandomState(6).rand(1024, 7)
    sb = api.synth_ifft_buff(full, None, 640, 1024, po["win_ISTFT"], 0.0, 5, 1.0)
    so = O.synth_ifft_buff(full, np.zeros_like(full), 640, 1024, po["win_ISTFT"], 0.0, 5, 1.0)
    assert rel_err(so, sb) < 1e-10


@pytest.mark.parametrize("l,gap", [(5, 3), (25, 3), (25, 1), (40, 5)])
def test_blk_sparse(api, O, l, gap):
    rs = np.random.RandomState(7)
    p = dict(api.default_p(), blk_gap=gap)
    po = dict(O.default_params(), blk_gap=gap)
    r_blk = rs.rand(513, 20)
    X, D = rs.rand(513) + 0.1, rs.rand(513) * (rs.rand(513) > 0.1)
    Q, rout = api.blk_sparse(X, D, r_blk, l, p)
    Qo, ro = O.blk_sparse(X, D, r_blk, l, po)
    assert np.allclose(rout, ro, rtol=1e-14, atol=0)
    assert np.allclose(Q, Qo, rtol=1e-12, atol=1e-15)


def test_per_hop_stream_matches_oracle_and_batch(api, O, bases, wavs, rng_inputs):
    """config 2 (latency path): init_buff + bnmf_sep_event_RT_IS16 hop by hop == oracle == batch path."""
    h_init, Ad = rng_inputs
    p = api.default_p()
    po = O.default_params()
    pcm = wavs["LM_in"][16000:16000 + 160 * 90]
    Bx, Bd = bases["B_DFT_x"], bases["B_DFT_d"]
    g = api.init_buff(Bx, Bd, Bx, Bd, p, Ad_blk_init=Ad)
    go = O.init_buff(Bx, Bd, Bx, Bd, po, Ad_blk_init=Ad)
    y = np.zeros(640)
    n_full = len(pcm) // 160
    frames = []
    for l in range(1, n_full + 4 + 1):
        if l <= n_full:
            y = np.concatenate([y[160:], pcm[(l - 1) * 160:l * 160].astype(float)])
        else:
            y = np.zeros(640)
        want_aux = l in (30, 60)
        xh, dh, xt, g = api.bnmf_sep_event_RT_IS16(y, l, g, p, h_init=h_init, nargout=3 if want_aux else 1)
        xho, dho, xto, go = O.bnmf_sep_event_RT_IS16(y, l, go, po, h_init=h_init, want_aux=want_aux)
        st = g["stats"]
        assert int(st[0]) == go.dbg["h_iters"], l
        assert bool(st[1]) == go.dbg["gated"] and int(st[2]) == go.dbg["R_a_up"] and int(st[3]) == go.dbg["w_iters"]
        assert np.max(np.abs(xt - xto)) <= 1e-6 * max(1.0, np.max(np.abs(xto))), l
        if want_aux:
            assert xh.shape == (1, 1, 640) and dh.shape == (1, 1, 640)
            assert np.max(np.abs(xh[0, 0] - xho[0])) <= 1e-6 * max(1.0, np.max(np.abs(xho)))
            assert np.max(np.abs(dh[0, 0] - dho[0])) <= 1e-6 * max(1.0, np.max(np.abs(dho)))
        frames.append(xt)
    for name, ref in (("B_DFT_d", go.B_DFT_d), ("Ad_blk", go.Ad_blk), ("lambda_d_blk", go.lambda_d_blk),
                      ("r_blk", go.r_blk), ("lambda_dav", go.lambda_dav), ("Xm_tilde", go.Xm_tilde), ("Ym", go.Ym)):
        assert rel_err(ref, g[name]) < 1e-9, name
    assert np.all(g["Yp"] == 0) and np.all(go.Yp == 0)      # the last flush hop is an all-zero frame
    # state write-back round trip (what a MATLAB caller does with B_D_u.mat)
    B = g["B_DFT_d"]
    g["B_DFT_d"] = B
    g["Ad_blk"] = g["Ad_blk"]
    assert np.array_equal(g["B_DFT_d"], B)
    assert rel_err(go.lambda_d_blk, g["lambda_d_blk"]) < 1e-9
    # same signal through the batch path
    out = api.enhance_batch([pcm], p, Bx, Bd, h_init=h_init, Ad_blk_init=Ad)[0]
    ola = np.zeros(640)
    ref = []
    for l, fr in enumerate(frames, start=1):
        if l > 3:
            ola = np.concatenate([ola[160:], np.zeros(160)]) + fr
            ref.append(O.to_int16(ola[:160]))
    assert np.abs(np.concatenate(ref).astype(int) - out.astype(int)).max() <= 1
    g.close()


def test_per_hop_stream_three_event_classes(api, O, bases, wavs, rng_inputs):
    """EVENT_NUM = 3, EVENT_RANK = [1 21 41] (initial_setting_Proposed_Techwin_201603_RT.m:40-49): the per-class
    reconstructions x_hat_i of bnmf_sep_event_RT_IS16.m:159-164,373-380 come back class by class."""
    h_init, Ad = rng_inputs
    over = dict(EVENT_NUM=3, EVENT_RANK=[1, 21, 41])
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    pcm = wavs["M04_in"][12000:12000 + 160 * 45]
    Bx, Bd = bases["B_DFT_x"], bases["B_DFT_d"]
    g = api.init_buff(Bx, Bd, Bx, Bd, p, Ad_blk_init=Ad)
    go = O.init_buff(Bx, Bd, Bx, Bd, po, Ad_blk_init=Ad)
    y = np.zeros(640)
    for l in range(1, len(pcm) // 160 + 1):
        y = np.concatenate([y[160:], pcm[(l - 1) * 160:l * 160].astype(float)])
        aux = l >= 30
        xh, dh, xt, g = api.bnmf_sep_event_RT_IS16(y, l, g, p, h_init=h_init, nargout=3 if aux else 1)
        xho, dho, xto, go = O.bnmf_sep_event_RT_IS16(y, l, go, po, h_init=h_init, want_aux=aux)
        assert int(g["stats"][0]) == go.dbg["h_iters"], l
        assert np.max(np.abs(xt - xto)) <= 1e-6 * max(1.0, np.max(np.abs(xto))), l
        if aux:
            assert xh.shape == (3, 1, 640)
            for i in range(3):
                assert np.max(np.abs(xh[i, 0] - xho[i])) <= 1e-6 * max(1.0, np.max(np.abs(xho))), (l, i)
    g.close() if hasattr(g, "close") else None


def test_per_hop_stream_semi_supervised(api, O, bases, wavs, rng_inputs):
    """p.basis_update_N = 1 through the per-hop entry (bnmf_sep_event_RT_IS16.m:125-127): the separation solve updates the
    noise half of the dictionary while it iterates; g.B_DFT_d itself only changes through the adaptation (:336)."""
    h_init, Ad = rng_inputs
    over = dict(basis_update_N=1, max_iter=25)
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    pcm = wavs["M03_in"][9000:9000 + 160 * 30]
    Bx, Bd = bases["B_DFT_x"], bases["B_DFT_d"]
    g = api.init_buff(Bx, Bd, Bx, Bd, p, Ad_blk_init=Ad)
    go = O.init_buff(Bx, Bd, Bx, Bd, po, Ad_blk_init=Ad)
    y = np.zeros(640)
    for l in range(1, len(pcm) // 160 + 1):
        y = np.concatenate([y[160:], pcm[(l - 1) * 160:l * 160].astype(float)])
        _, _, xt, g = api.bnmf_sep_event_RT_IS16(y, l, g, p, h_init=h_init, nargout=1)
        _, _, xto, go = O.bnmf_sep_event_RT_IS16(y, l, go, po, h_init=h_init)
        assert int(g["stats"][0]) == go.dbg["h_iters"], l
        assert np.max(np.abs(xt - xto)) <= 1e-6 * max(1.0, np.max(np.abs(xto))), l
    assert rel_err(go.B_DFT_d, g["B_DFT_d"]) < 1e-9
    g.close() if hasattr(g, "close") else None


@pytest.mark.parametrize("mel", [False, True], ids=["run_basis_DNMF", "run_basis_DNMF_Mel"])
def test_dnmf_basis_retraining_matches_oracle(api, O, wavs, mel):
    """run_basis_DNMF.m / run_basis_DNMF_Mel.m: STFT of clean, noise and mixture, activations of
    the mixture with the dictionary fixed, then W-only updates of the two halves -- every step on the GPU."""
    rs = np.random.RandomState(9)
    x = wavs["M03_in"][8000:8000 + 16000].astype(np.float64)
    d = wavs["M04_in"][3000:3000 + 15000].astype(np.float64) * 0.5        # shorter: both are cut to its length (:5-9)
    R_x, R_d = 10, 6
    over = dict(R_x=R_x, R_d=R_d, max_iter=12, conv_eps=1e-4)
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    F = 64 if mel else 513
    B = rs.rand(F, R_x + R_d) + 0.05
    h0 = rs.rand(R_x + R_d, 200)
    rand = lambda m, n: h0[:m, :n].copy()
    B_hat = api.run_basis_DNMF(x, d, B, p, rand=rand, mel=mel)
    B_ref = O.run_basis_dnmf(x, d, B, po, rand=rand, mel=mel)
    assert B_hat.shape == B_ref.shape == (F, R_x + R_d)
    assert rel_err(B_ref, B_hat) < TOL, rel_err(B_ref, B_hat)
    assert np.allclose(np.linalg.norm(B_hat, axis=0), 1.0, atol=1e-9)     # sparse_nmf.m:242
    assert rel_err(B, B_hat) > 1e-2                                        # it did move


def test_tf_features(api, O):
    rs = np.random.RandomState(2)
    S = rs.rand(513, 37)
    M = O.mel_matrix(16000, 64, 1024, 1, 8000)
    np.testing.assert_allclose(api.tf_features(S, 2.0, 1e-9), S ** 2 + 1e-9, rtol=1e-15)
    np.testing.assert_allclose(api.tf_features(S, 1.5, 1e-9), S ** 1.5 + 1e-9, rtol=1e-13)
    np.testing.assert_allclose(api.tf_features(S, 2.0, 1e-9, M), M.T @ (S ** 2 + 1e-9), rtol=1e-12)


@pytest.mark.parametrize("variant_c,with_a", [(False, False), (True, False), (False, True)])
def test_gist_ntf_matches_oracle(api, O, variant_c, with_a):
    """GIST_NTF / GIST_NTF_C (6-channel KL tensor factorisation with fixed dictionary, channel gains updated)."""
    rs = np.random.RandomState(13)
    Ch, N, M, K = 6, 129, 23, 17
    B = rs.rand(N, K) + 0.01
    Ct = rs.rand(Ch, K)
    S = np.einsum('hk,nk,mk->hnm', Ct, B, rs.gamma(1.0, 1.0, (M, K))) + 1e-6
    C0 = rs.rand(Ch, K)
    rand = lambda a, b: C0[:a, :b].copy()
    A = rs.rand(M, K) + 0.1 if with_a else None
    p = dict(max_iter=25, conv_eps=1e-4, sparsity=0.5, nonzerofloor=1e-9, cost_check=1)
    Cg, Ag, og = api.GIST_NTF(p, B, S, rand=rand, A=A, variant_c=variant_c)
    Cr, Ar, orf = O.gist_ntf(p, B, S, rand=rand, A=A, variant_c=variant_c)
    assert og["iters"] == orf["iters"]
    assert rel_err(Cr, Cg) < 1e-9
    np.testing.assert_allclose(og["cost"], orf["cost"], rtol=1e-10)
    assert np.all(np.diff(og["cost"]) <= 1e-9 * og["cost"][:-1])        # KL + L1 objective does not increase
    assert Ag.shape == (M, K)
    # GIST_NTF_C with p.cost_check = 0 runs max_iter iterations and reports no objective
    if variant_c:
        Cg2, _, og2 = api.GIST_NTF(dict(p, cost_check=0, max_iter=5), B, S, rand=rand, variant_c=True)
        Cr2, _, or2 = O.gist_ntf(dict(p, cost_check=0, max_iter=5), B, S, rand=rand, variant_c=True)
        assert og2["iters"] == or2["iters"] == 5 and og2["cost"].size == 0
        assert rel_err(Cr2, Cg2) < 1e-9


def test_stream_fixed_columns_follow_the_mel_slot_after_updates(api, O, bases, wavs, rng_inputs):
    """bnmf_sep_event_RT_IS16.m:328 [sic]: even in DFT mode the never-updated noise atoms are re-assembled from
    g.B_Mel_d(:, R_a+1:end) on every update.  With a B_Mel_d slot that DIFFERS from B_DFT_d the fixed columns therefore
    switch to B_Mel_d's at the first update and must stay there over later updates (both ping-pong buffers)."""
    h_init, Ad = rng_inputs
    p = api.default_p()
    po = O.default_params()
    rs = np.random.RandomState(17)
    Bx, Bd = bases["B_DFT_x"], bases["B_DFT_d"]
    Bmel = Bd * np.exp(0.2 * rs.randn(*Bd.shape))            # a different "B_Mel_d" slot of the same shape
    g = api.init_buff(Bx, Bmel, Bx, Bd, p, Ad_blk_init=Ad)
    go = O.init_buff(Bx, Bmel, Bx, Bd, po, Ad_blk_init=Ad)
    pcm = wavs["M03_in"][8000:8000 + 160 * 60]
    y = np.zeros(640)
    updates = 0
    for l in range(1, 61):
        y = np.concatenate([y[160:], pcm[(l - 1) * 160:l * 160].astype(float)])
        _, _, xt, g = api.bnmf_sep_event_RT_IS16(y, l, g, p, h_init=h_init, nargout=1)
        _, _, xto, go = O.bnmf_sep_event_RT_IS16(y, l, go, po, h_init=h_init)
        updates += int(go.dbg["w_iters"] > 0)
        assert int(g["stats"][0]) == go.dbg["h_iters"], l
        assert np.max(np.abs(xt - xto)) <= 1e-6 * max(1.0, np.max(np.abs(xto))), l
    assert updates >= 3
    assert rel_err(go.B_DFT_d, g["B_DFT_d"]) < 1e-9
    assert rel_err(Bmel[:, 50:], g["B_DFT_d"][:, 50:]) < 1e-12
    g.close()


def test_per_hop_stream_mel_mode(api, O, bases, wavs, rng_inputs):
    """p.B_sep_mode = 'Mel' through init_buff + bnmf_sep_event_RT_IS16 hop by hop (the latency entry): separation and
    adaptation on the 64-band Mel dictionaries, everything else in the DFT domain, exactly like the batch entry."""
    h_init, Ad = rng_inputs
    over = dict(B_sep_mode="Mel", MelConv=1)
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    Bx, Bd, Bmx, Bmd = bases["B_DFT_x"], bases["B_DFT_d"], bases["B_Mel_x"], bases["B_Mel_d"]
    g = api.init_buff(Bmx, Bmd, Bx, Bd, p, Ad_blk_init=Ad)
    go = O.init_buff(Bmx, Bmd, Bx, Bd, po, Ad_blk_init=Ad)
    pcm = wavs["M03_in"][6000:6000 + 160 * 70]
    y = np.zeros(640)
    updates = 0
    for l in range(1, 71):
        y = np.concatenate([y[160:], pcm[(l - 1) * 160:l * 160].astype(float)])
        aux = l in (20, 50)
        xh, dh, xt, g = api.bnmf_sep_event_RT_IS16(y, l, g, p, h_init=h_init, nargout=3 if aux else 1)
        xho, dho, xto, go = O.bnmf_sep_event_RT_IS16(y, l, go, po, h_init=h_init, want_aux=aux)
        updates += int(go.dbg["w_iters"] > 0)
        st = g["stats"]
        assert int(st[0]) == go.dbg["h_iters"] and int(st[3]) == go.dbg["w_iters"], l
        assert np.max(np.abs(xt - xto)) <= 1e-6 * max(1.0, np.max(np.abs(xto))), l
        if aux:
            assert np.max(np.abs(xh[0, 0] - xho[0])) <= 1e-6 * max(1.0, np.max(np.abs(xho)))
            assert np.max(np.abs(dh[0, 0] - dho[0])) <= 1e-6 * max(1.0, np.max(np.abs(dho)))
    assert updates > 0
    assert g["B_Mel_d"].shape == (64, 100) and rel_err(go.B_Mel_d, g["B_Mel_d"]) < 1e-9
    assert rel_err(go.lambda_d_blk, g["lambda_d_blk"]) < 1e-9
    g.close()
