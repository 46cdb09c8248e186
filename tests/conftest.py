import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(GOLDEN / f'{name}.npz', allow_pickle=False)
    return load


def conw_inputs(seed, n, d, n_clients):
    """Same generator as tests/golden/make_golden.py::conw_inputs (numpy PCG64, stable across platforms)."""
    rng = np.random.default_rng(seed)

    def unit_np(x):
        return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)

    g_img = unit_np(rng.standard_normal((n, d)))
    g_txt = unit_np(0.7 * g_img + 0.5 * unit_np(rng.standard_normal((n, d))))
    i_vecs = [unit_np(g_img + (0.3 + 0.35 * c) * unit_np(rng.standard_normal((n, d)))) for c in range(n_clients)]
    t_vecs = [unit_np(g_txt + (0.3 + 0.35 * c) * unit_np(rng.standard_normal((n, d)))) for c in range(n_clients)]
    return g_img, g_txt, i_vecs, t_vecs
