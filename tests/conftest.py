import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_npz_groups(name):
    """Fixture files store 'case/key' entries; returns ({case: {key: array}}, {top-level key: array})."""
    z = np.load(os.path.join(GOLDEN, name))
    groups, top = {}, {}
    for k in z.files:
        if '/' in k:
            c, kk = k.split('/', 1)
            groups.setdefault(c, {})[kk] = z[k]
        else:
            top[k] = z[k]
    return groups, top


@pytest.fixture(scope='session')
def golden_weights():
    from diffndm_b200.weights import DynamicsConfig, random_init, weights_checksum
    _, top = load_npz_groups('forward.npz')
    W = random_init(DynamicsConfig(), int(top['weight_seed']), float(top['coord_head_gain']))
    assert abs(weights_checksum(W) - float(top['weights_checksum'])) < 1e-6 * max(1.0, abs(float(top['weights_checksum']))), \
        'random_init drifted from the weights the golden fixtures were generated with'
    return W
