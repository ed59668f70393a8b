import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

CONFIG_DIR = ROOT / "tests" / "data" / "model_configs"
MODELS = {
    "fast": "dna_r10.4.1_e8.2_400bps_fast@v5.0.0",
    "hac": "dna_r10.4.1_e8.2_400bps_hac@v5.0.0",
    "sup": "dna_r10.4.1_e8.2_400bps_sup@v5.0.0",
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def model_dir(kind: str) -> pathlib.Path:
    return CONFIG_DIR / MODELS[kind]


def synthetic_scores(N, T, state_len, seed, scale=2.0, dtype=np.float16):
    """Random CRF transition scores with some temporal structure so the beam has real choices."""
    rng = np.random.default_rng(seed)
    C = 4 ** (state_len + 1)
    base = rng.standard_normal((N, T, C)).astype(np.float32) * scale
    # favour a random "true" path so stays and steps alternate
    return np.clip(base, -6, 6).astype(dtype)


@pytest.fixture(scope="session")
def crf_oracle():
    from oracle.oracle import CrfOracle, build
    build(ref=False)
    return CrfOracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.oracle import Reference, reference_available
    if not reference_available():
        pytest.skip("oracle/_ref/libdorado_ref.so not built (needs /root/reference)")
    return Reference()
