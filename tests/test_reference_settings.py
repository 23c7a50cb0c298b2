"""The reference's own settings files (tests/golden/reference_settings.json, written from settings/*.m and
settings/bak_IS16_results/*.m by tests/golden/make_settings_fixture.py) as parameter sets of the drop-in.

CPU part: the fixture is what the library and the oracle call "shipped defaults".  GPU part: every file that selects the
SNMF path runs through the C ABI and must reproduce the oracle hop by hop (iteration counts, gates, activations, PCM)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, rel_err, snr_db

SETTINGS = json.loads((GOLDEN / "reference_settings.json").read_text())
SHIPPED = "settings/initial_setting_SNMF_NAT.m"
SPEC_TOL = 1e-3
WAVE_SNR_DB = 40.0


def overrides(rec):
    return {k: v for k, v in rec.items() if not k.startswith("_") and k != "NMF_algorithm"}


def test_fixture_covers_every_settings_file():
    assert len(SETTINGS) == 9 and SHIPPED in SETTINGS
    assert all(rec["_win_is_sqrt_hann_periodic"] for rec in SETTINGS.values())
    # the kl divergence and the Mel -> DFT conversion are the only ones any file selects
    assert {rec["cf"] for rec in SETTINGS.values()} == {"kl"}
    assert {rec["MelConv"] for rec in SETTINGS.values()} == {1}


def test_shipped_settings_file_is_the_default_parameter_set():
    """snmfnat_params_default (C ABI) and oracle.default_params restate settings/initial_setting_SNMF_NAT.m."""
    from oracle import snmf_oracle as O
    from se_snmf_nat_b200 import api
    rec = overrides(SETTINGS[SHIPPED])
    lib_p, ora_p = api.default_p(), O.default_params()
    for k, v in rec.items():
        for name, p in (("library", lib_p), ("oracle", ora_p)):
            if k not in p:
                continue
            got = p[k]
            if isinstance(v, list):
                assert list(got)[:len(v)] == v, (name, k)
            elif isinstance(v, str):
                assert str(got) == v, (name, k)
            else:
                assert float(got) == pytest.approx(float(v), rel=1e-15), (name, k)
    # every numeric field of the fixture is known to the parameter struct
    from se_snmf_nat_b200._lib import Params
    known = {n for n, _ in Params._fields_} | {"random_seed", "useGPU", "B_sep_mode", "cf", "ENHANCE_METHOD"}
    assert set(rec) <= known, set(rec) - known


def _grow(B, cols, seed):
    rs = np.random.RandomState(seed)
    reps = [B[:, rs.permutation(B.shape[1])] * np.exp(0.3 * rs.randn(*B.shape)) for _ in range((cols + B.shape[1] - 1) // B.shape[1])]
    return np.concatenate([B] + reps, axis=1)[:, :cols].copy()


SNMF_FILES = sorted(k for k, rec in SETTINGS.items() if rec["NMF_algorithm"] == "SNMF")


@pytest.mark.gpu
@pytest.mark.parametrize("name", SNMF_FILES, ids=[k.split("initial_setting_")[1][:-2] for k in SNMF_FILES])
def test_settings_file_runs_like_the_oracle(name, bases, wavs):
    from oracle import snmf_oracle as O
    from se_snmf_nat_b200 import api
    over = overrides(SETTINGS[name])
    if "R_a" not in over:   # files older than the adaptation (adapt_train_N = 0): only the shape of Ad_blk depends on it
        over["R_a"] = min(50, over["R_d"])
    p = dict(api.default_p(), **over)
    po = dict(O.default_params(), **over)
    R_x, R_d = p["R_x"], p["R_d"]
    Bx = bases["B_DFT_x"] if R_x == 100 else _grow(bases["B_DFT_x"], R_x, 1)
    Bd = bases["B_DFT_d"] if R_d == 100 else _grow(bases["B_DFT_d"], R_d, 3)
    h_init = O.park_miller(R_x + R_d, 1)
    Ad = np.random.RandomState(17).rand(p["R_a"], p["m_a"])
    ctx = api.get_context(0)
    if name.endswith("SNMF_Techwin_201603_RT.m"):
        # EVENT_RANK = [1 21 41] with R_x = 20: B_DFT_x(:, 21:40) is an index error in the reference (:160); refused here
        with pytest.raises(api.SnmfnatError) as e:
            api.Batch(ctx, p, Bx, Bd, [4000], h_init, Ad)
        assert "EVENT_RANK" in str(e.value)
        return
    hops = 24 if R_x + R_d > 400 else 60
    pcms = [wavs["M03_in"][8000:8000 + 160 * hops], wavs["M04_in"][3000:3000 + 160 * (hops - 9) + 31]]
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
        assert np.array_equal(b.trace(i, "gated").astype(int), np.array([int(t["gated"]) for t in tr])), i
        A = b.trace(i, "A")
        assert max(rel_err(tr[k]["A"], A[k]) for k in range(len(tr))) <= SPEC_TOL
        assert snr_db(ref, outs[i]) >= WAVE_SNR_DB
        assert np.abs(outs[i].astype(int) - ref.astype(int)).max() <= 1, i
    b.close()
