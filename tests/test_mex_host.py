"""The MEX boundary, EXECUTED: every gateway of mex/ is linked against a small working host (mex/mex_host.cpp implements
the MEX C API of mex/mex_shim.h) and libsnmfnat.so, then called through mexFunction exactly as MATLAB / Octave would call
it: struct p, struct g, column-major doubles, rand drawn by the host through mexCallMATLAB.  CPU tests
cover linking, argument errors and the one gateway without device work; the GPU tests compare with the oracle."""
import ctypes as C
import wave

import numpy as np
import pytest

import mexhost
from conftest import rel_err, snr_db

GATEWAYS = ["sparse_nmf", "snmf_mdi", "snmf_mdi_Sm", "DNMF_adapt", "stft_fft", "synth_ifft_buff", "blk_sparse",
            "bnmf_sep_event_RT_IS16", "snmfnat_enhance_files", "mel_matrix", "GIST_NTF"]


@pytest.fixture(scope="module")
def host():
    from se_snmf_nat_b200 import build as b
    b.build()                     # libsnmfnat.so must exist before the gateways link against it
    mexhost.build()
    return mexhost.Host()


def test_every_gateway_links_and_exports_mexFunction(host):
    for g in GATEWAYS:
        lib = host.gateway(g)
        assert C.cast(lib.mexFunction, C.c_void_p).value, g


def test_usage_errors_come_back_as_mex_errors(host):
    with pytest.raises(RuntimeError, match="snmfnat:usage"):
        host.call("sparse_nmf", 3, np.ones((4, 3)))                         # p missing
    with pytest.raises(RuntimeError, match="snmfnat:usage"):
        host.call("mel_matrix", 1, 16000.0)
    with pytest.raises(RuntimeError, match="cost_check is required"):      # sparse_nmf.m:260 reads p.cost_check unguarded
        host.call("sparse_nmf", 2, np.ones((4, 3)), dict(r=2.0, max_iter=3.0, random_seed=1.0))
    assert host.lib.mexhost_live_arrays() == 0                             # nothing leaked on the error paths


def test_mel_matrix_gateway_runs_end_to_end(host):
    from oracle import snmf_oracle as O
    M, = host.call("mel_matrix", 1, 16000.0, 64.0, 1024.0)
    assert M.shape == (513, 64) and np.array_equal(M, O.mel_matrix(16000, 64, 1024, 1.0, None))
    M2, = host.call("mel_matrix", 1, 8000.0, 24.0, 256.0, 1.0, 3800.0)
    assert np.array_equal(M2, O.mel_matrix(8000, 24, 256, 1.0, 3800.0))


@pytest.mark.gpu
def test_sparse_nmf_gateway_matches_oracle(host):
    """[w, h, objective] = sparse_nmf(v, p) with the random init drawn by the host's rand after rand('seed', 1)."""
    from oracle import snmf_oracle as O
    rs = np.random.RandomState(4)
    m, n, r = 65, 40, 6
    V = rs.gamma(1.0, 1.0, (m, n)) + 1e-3
    p = dict(r=float(r), max_iter=30.0, random_seed=1.0, cost_check=1.0, sparsity=0.5, conv_eps=1e-4, cf="kl")
    w, h, obj = host.call("sparse_nmf", 3, V, p)
    seq = O.park_miller(m * r + r * n, 1)                                   # rand(m, r) then rand(r, n), column-major
    w0 = seq[:m * r].reshape(m, r, order="F")
    h0 = seq[m * r:].reshape(r, n, order="F")
    wr, hr, objr = O.sparse_nmf(V, init_w=w0, init_h=h0, max_iter=30, sparsity=0.5, conv_eps=1e-4, cf="kl", cost_check=True)
    assert rel_err(wr, w) < 1e-9 and rel_err(hr, h) < 1e-9
    assert obj["cost"].shape[1] == len(objr["cost"])
    np.testing.assert_allclose(obj["cost"].ravel(), objr["cost"], rtol=1e-9)


@pytest.mark.gpu
def test_sparse_nmf_gateway_reaches_the_tensor_core_training_path(host):
    """run_basis_train.m:84-88 through the drop-in: every atom updated, KL, n >= 16384 and p.useGPU ~= 0 -> snmfnat_train_*
    (tf32 tensor cores).  Same call with p.useGPU = 0 stays on the float64 kernels; both within 1e-3 of the oracle."""
    from oracle import snmf_oracle as O
    rs = np.random.RandomState(9)
    F, K, T = 513, 32, 16384
    Wt = np.abs(rs.randn(F, 24))
    V = (Wt / np.linalg.norm(Wt, axis=0)) @ rs.gamma(0.3, 1.0, (24, T)) + 1e-9
    V = V.astype(np.float32).astype(np.float64)
    w0 = V[:, rs.choice(T, K, replace=False)]
    h0 = rs.rand(K, T).astype(np.float32).astype(np.float64)
    p = dict(init_w=w0, init_h=h0, max_iter=8.0, random_seed=1.0, cost_check=1.0, sparsity=5.0, conv_eps=0.0, cf="kl",
             w_update_ind=np.ones(K, bool), h_update_ind=np.ones(K, bool), useGPU=1.0)
    wr, hr, objr = O.sparse_nmf(V, init_w=w0, init_h=h0, max_iter=8, sparsity=5.0, conv_eps=0.0, cf="kl", cost_check=True,
                                w_update_ind=np.ones(K, bool), h_update_ind=np.ones(K, bool))
    launches = []
    for use in (1.0, 0.0):
        p["useGPU"] = use
        w, h, obj = host.call("sparse_nmf", 3, V, p)
        assert rel_err(wr, w) < 1e-3 and rel_err(hr, h) < 1e-3, use
        np.testing.assert_allclose(obj["cost"].ravel(), objr["cost"], rtol=1e-3)
        launches.append(rel_err(wr, w))
    assert launches[0] > 1e-9 > launches[1], launches    # tf32 path really ran for useGPU = 1, float64 for useGPU = 0


@pytest.mark.gpu
def test_bnmf_sep_event_gateway_hop_loop_matches_oracle(host, bases, wavs, rng_inputs):
    """filewise_run_IS16.m:86-169 with the gateway in place of bnmf_sep_event_RT_IS16.m: struct g from init_buff goes in
    and out by value, rand('seed',1); rand(R,1) is drawn through mexCallMATLAB on every hop; a second file that re-uses
    the struct (l == 1 again) releases the first file's device stream."""
    from oracle import snmf_oracle as O
    from se_snmf_nat_b200 import api
    h_init, Ad = rng_inputs
    po = O.default_params()
    p = mexhost.matlab_p(api.default_p())
    Bx, Bd = bases["B_DFT_x"], bases["B_DFT_d"]
    g0 = dict(B_Mel_x=Bx, B_Mel_d=Bd, B_DFT_x=Bx, B_DFT_d=Bd, Ad_blk=Ad, update_switch=1.0)
    L = host.lib
    for file_no, off in enumerate((16000, 40000)):
        pcm = wavs["LM_in"][off:off + 160 * 40]
        go = O.init_buff(Bx, Bd, Bx, Bd, po, Ad_blk_init=Ad)
        if file_no == 0:
            g = host.mx(g0)
        else:
            # the next file: a fresh init_buff struct, but still carrying the handle field of the previous file (what a
            # MATLAB loop that rebuilds g field by field would hand over): the gateway must release that stream
            old = host.py(g)["snmfnat_handle"]
            L.mxDestroyArray(g)
            g = host.mx(g0)
            hnd = L.mxCreateNumericArray(2, (C.c_size_t * 2)(1, 1), mexhost.UINT64, 0)
            C.memmove(L.mxGetData(hnd), np.asarray(old, dtype=np.uint64).ctypes.data, 8)
            L.mxAddField(g, b"snmfnat_handle")
            L.mxSetField(g, 0, b"snmfnat_handle", hnd)
        y = np.zeros(640)
        for l in range(1, 41):
            y = np.concatenate([y[160:], pcm[(l - 1) * 160:l * 160].astype(float)])
            outs = host.call("bnmf_sep_event_RT_IS16", 4, y.reshape(1, -1), float(l), g, p, keep=True)
            xt = host.py(outs[2])
            for o in outs[:3]:
                L.mxDestroyArray(o)
            L.mxDestroyArray(g)
            g = outs[3]
            _, _, xto, go = O.bnmf_sep_event_RT_IS16(y, l, go, po, h_init=h_init)
            assert xt.shape == (1, 640)
            assert np.max(np.abs(xt.ravel() - xto)) <= 1e-6 * max(1.0, np.max(np.abs(xto))), (file_no, l)
        gd = host.py(g)
        assert rel_err(go.B_DFT_d, gd["B_DFT_d"]) < 1e-9          # the field the reference's callers read back
        assert int(gd["snmfnat_handle"].ravel()[0]) == file_no + 1
    L.mxDestroyArray(g)
    host.lib.mexhost_shutdown()                                    # mexAtExit: streams and context released


def _write_wav(path, pcm):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(np.asarray(pcm, dtype="<i2").tobytes())


def _read_wav(path):
    with wave.open(str(path), "rb") as w:
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")


@pytest.mark.gpu
def test_enhance_files_gateway_one_and_many_gpus(host, bases, wavs, tmp_path):
    """snmfnat_enhance_files(paths_in, paths_out, B_DFT_x, B_DFT_d, p): WAV in, WAV out, rand through the host in the order
    of init_buff.m:37-38; with p.gpus = [0 1] the files are split over two devices (snmfnat_enhance_batch_multi) and the
    result must not change."""
    import torch
    from oracle import snmf_oracle as O
    from se_snmf_nat_b200 import api
    po = O.default_params()
    m04 = wavs["M04_in"]
    pcms = [m04[o:o + n] for o, n in [(0, 160 * 30 + 7), (9000, 160 * 45), (20000, 160 * 22), (33000, 160 * 38 + 90)]]
    ins, outs = [], []
    for i, x in enumerate(pcms):
        _write_wav(tmp_path / f"in{i}.wav", x)
        ins.append(str(tmp_path / f"in{i}.wav"))
        outs.append(str(tmp_path / f"out{i}.wav"))
    rs = np.random.RandomState(31)
    stream = np.concatenate([np.concatenate([rs.rand(100), rs.rand(50 * 100)]) for _ in pcms])   # per file: A_d, then Ad_blk
    p = mexhost.matlab_p(api.default_p())
    Bx, Bd = bases["B_DFT_x"], bases["B_DFT_d"]
    h_init = O.park_miller(200, 1)
    configs = [None] + ([[0.0, 1.0]] if torch.cuda.device_count() >= 2 else [])
    for gpus in configs:
        q = dict(p)
        if gpus:
            q["gpus"] = np.array(gpus)
        # the gateway seeds first (rand('seed',1): Park-Miller from then on), so the per-file draws are Park-Miller too
        host.set_rand_stream(stream)
        host.call("snmfnat_enhance_files", 0, ins, outs, Bx, Bd, q)
        seq = O.park_miller(200 + len(pcms) * (100 + 5000), 1)
        pos = 200
        for i, x in enumerate(pcms):
            ad = seq[pos + 100: pos + 5100].reshape(50, 100, order="F")
            pos += 5100
            ref, _ = O.enhance_utterance(x, po, Bx, Bd, h_init=h_init, Ad_blk_init=ad)
            got = _read_wav(outs[i])
            want = api.pcm2wav_samples(ref)                 # src/pcm2wav.m: wavwrite re-quantisation of the raw PCM
            assert len(got) == len(want)
            assert np.abs(got.astype(int) - want.astype(int)).max() <= 1, (gpus, i)
