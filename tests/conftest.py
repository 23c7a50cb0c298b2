import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def bases():
    z = np.load(GOLDEN / "bases.npz")
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def wavs():
    z = np.load(GOLDEN / "wavs.npz")
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def rng_inputs():
    z = np.load(GOLDEN / "rng_seed1.npz")
    return z["h_init"], z["Ad_blk"]


@pytest.fixture(scope="session")
def m03_oracle():
    z = np.load(GOLDEN / "M03_oracle.npz")
    return {k: z[k] for k in z.files}


def snr_db(ref, x):
    n = min(len(ref), len(x))
    ref = np.asarray(ref[:n], dtype=np.float64)
    x = np.asarray(x[:n], dtype=np.float64)
    err = np.sum((ref - x) ** 2)
    if err == 0:
        return np.inf
    return 10 * np.log10(np.sum(ref ** 2) / err)


def rel_err(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return float(np.linalg.norm(ref - x) / max(np.linalg.norm(ref), 1e-300))
