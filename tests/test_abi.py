"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol include/snmfnat.h
declares; the host-side parameter marshalling matches the shipped settings.  No compute calls (no GPU here)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from se_snmf_nat_b200 import build, _lib
    build.build()
    return _lib.load()


def header_symbols():
    txt = (ROOT / "include" / "snmfnat.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(snmfnat_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    from se_snmf_nat_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/snmfnat.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.snmfnat_version() == 100


def test_params_default_matches_shipped_settings(lib):
    from se_snmf_nat_b200 import api
    from oracle import snmf_oracle as O
    p = api.default_p()
    po = O.default_params()
    for k in ("framelength", "frameshift", "fftlength", "delay", "R_x", "R_d", "R_a", "m_a", "init_N_len",
              "P_len_k", "P_len_l", "blk_gap", "DCbin", "DCbin_back", "max_iter", "overlapscale", "pow",
              "nonzerofloor", "overlap_m_a", "Ar_up", "alpha_p", "preemph", "sparsity", "conv_eps", "alpha_eta",
              "alpha_d", "beta", "beta_max"):
        assert p[k] == po[k], k
    assert np.allclose(p["win_STFT"], po["win_STFT"], atol=0, rtol=0)
    ps = api.params_struct(dict(p, ENHANCE_METHOD="Wiener", EVENT_NUM=3, EVENT_RANK=[1, 21, 41], cf="is"))
    assert ps.ENHANCE_METHOD == 1 and list(ps.EVENT_RANK)[:3] == [1, 21, 41] and ps.cf == 0


def test_no_device_fails_loudly(lib):
    """On a box without a GPU context creation must fail with ENODEVICE -- never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from se_snmf_nat_b200 import api
    with pytest.raises(api.SnmfnatError) as e:
        api.Context(0)
    assert e.value.code == -2
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    for f in (ROOT / "se_snmf_nat_b200").rglob("*.py"):
        assert "oracle" not in f.read_text().replace("no oracle", ""), f


def test_mex_gateways_compile_against_shim():
    """No MATLAB/Octave here: the gateways are syntax- and ABI-checked against mex/mex_shim.h."""
    import subprocess
    r = subprocess.run(["make", "-C", str(ROOT / "mex"), "check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all gateways compile" in r.stdout


def test_mel_matrix_matches_the_oracle():
    """snmfnat_mel_matrix (src/mel_matrix.m) is host code of the library: checked here without a GPU."""
    import numpy as np
    from se_snmf_nat_b200 import api
    from oracle import snmf_oracle as O
    for args in [(16000, 64, 1024, 1.0, None), (16000, 40, 512, 1.0, None), (8000, 24, 256, 1.0, 3800.0),
                 (16000, 64, 1024, 1.1, None)]:
        a = api.mel_matrix(*args)
        b = O.mel_matrix(*args)
        assert a.shape == b.shape and np.array_equal(a, b)
    M = api.mel_matrix(16000, 64, 1024)
    assert M.min() == 0.0 and M.max() == 1.0 and np.all(M.sum(0) > 0)


def test_pcm2wav_requantisation():
    """src/pcm2wav.m:9-10 + wavwrite: round(s / 32767 * 32768), half away from zero, clipped to int16."""
    import numpy as np
    from se_snmf_nat_b200 import api
    s = np.array([0, 1, -1, 100, 16383, 16384, -16384, 20000, -20000, 32766, 32767, -32767, -32768], dtype=np.int16)
    want = [0, 1, -1, 100, 16383, 16385, -16385, 20001, -20001, 32767, 32767, -32768, -32768]
    assert api.pcm2wav_samples(s).tolist() == want


def test_basis_mat_round_trip(tmp_path, bases):
    """run_basis_train.m:136-138: the four variables a training run saves load back unchanged, and in the layout the
    reference's shipped basis files have (B_DFT_sub 513 x R, B_Mel_sub 64 x R)."""
    import numpy as np
    from se_snmf_nat_b200 import api
    f = tmp_path / "R_100.mat"
    A = np.random.RandomState(0).rand(100, 37)
    api.save_basis_mat(str(f), bases["B_DFT_x"], bases["B_Mel_x"], A, None)
    m = api.load_basis_mat(str(f))
    assert set(m) == {"B_DFT_sub", "B_Mel_sub", "A_DFT_sub", "A_Mel_sub"}
    assert np.array_equal(m["B_DFT_sub"], bases["B_DFT_x"]) and np.array_equal(m["B_Mel_sub"], bases["B_Mel_x"])
    assert np.array_equal(m["A_DFT_sub"], A) and m["A_Mel_sub"].shape == (1, 1)
    assert f.read_bytes()[:19] == b"MATLAB 5.0 MAT-file"
