"""Loading helpers of tests/golden/parity_r2.npz shared by the CPU (oracle) and GPU (engine) tests."""
import os
import sys

import numpy as np

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from guidance_common import NoiseStream, stress_weights  # noqa: E402


def load_r2():
    z = np.load(os.path.join(GOLDEN, 'parity_r2.npz'))
    groups, top = {}, {}
    for k in z.files:
        if '/' in k:
            c, kk = k.split('/', 1)
            groups.setdefault(c, {})[kk] = z[k]
        else:
            top[k] = z[k]
    return groups, top


R2, R2_TOP = load_r2()
FWD = sorted(k for k in R2 if k.startswith('fwd_'))
TRAJ = sorted(k for k in R2 if k.startswith('traj_'))
INP = sorted(k for k in R2 if k.startswith('inp_'))


def weights_for(name, base):
    """The weight table a forward case was generated with (regenerated, never stored; guarded by the case's checksum)."""
    from diffndm_b200.weights import DynamicsConfig, random_init, weights_checksum
    if 'untied' in name:
        W = random_init(DynamicsConfig(), int(R2_TOP['weight_seed']), float(R2_TOP['coord_head_gain']), untie_heads=True)
    elif 'r2stress' in name:
        W = stress_weights(base)
    else:
        W = base
    ref = float(R2[name]['weights_checksum'])
    assert abs(weights_checksum(W) - ref) < 1e-6 * max(1.0, abs(ref)), name
    return W


def traj_draws(c):
    """z_T draw, one per step, final head -- regenerated from the seed."""
    s = NoiseStream(int(c['noise_seed']))
    n_l = int(np.sum(c['sizes']))
    return [s.draw(n_l, c['x0_rel'].shape[1]) for _ in range(int(c['timesteps']) + 2)]


def inpaint_draws(c):
    s = NoiseStream(int(c['noise_seed']))
    return [s.draw(len(c['lig_mask']), c['x0_rel'].shape[1]) for _ in range(int(c['n_draws']))]


def batch_of(c):
    B, n_p = len(c['sizes']), len(c['pocket_x'])
    return B, n_p, np.repeat(np.arange(B, dtype=np.int64), c['sizes']), np.repeat(np.arange(B, dtype=np.int64), n_p)
