"""Host-side pieces of the product that run without a GPU, against the oracle (itself pinned to the reference fixtures):
the noise schedule table and the per-step coefficients the fused sampler-step kernel consumes."""
import numpy as np
import torch

from oracle import egnn_oracle as O


def test_schedule_table_matches_oracle():
    from diffndm_b200.sampler import polynomial_gamma
    for T, prec, power in ((500, 5e-4, 2.0), (100, 1e-4, 2.0), (50, 5e-4, 3.0)):
        g = polynomial_gamma(T, prec, power)
        ref = O.gamma_table(T, prec, power)
        assert g.dtype == torch.float32 and tuple(g.shape) == (T + 1,)
        assert np.abs(g.numpy() - ref).max() < 2e-6 * np.abs(ref).max()


def test_step_coefficients_match_oracle():
    """(1/alpha_ts, sigma2_ts/alpha_ts/sigma_t, sigma_ts*sigma_s/sigma_t) of conditional_model.py:486-529."""
    from diffndm_b200.sampler import ConditionalSampler
    g = torch.from_numpy(O.gamma_table(500, 5e-4, 2.0))
    coef = ConditionalSampler.step_coefficients(None, g[:-1], g[1:]).numpy()
    sc = O.step_scalars(g[:-1].numpy(), g[1:].numpy())
    ref = np.stack([1.0 / sc['alpha_ts'], sc['sigma2_ts'] / sc['alpha_ts'] / sc['sigma_t'],
                    sc['sigma_ts'] * sc['sigma_s'] / sc['sigma_t']], 1)
    assert coef.shape == (500, 3)
    assert np.abs(coef - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())
    assert coef[:, 1].max() < 0.18                     # c_eps <= 0.177 over polynomial_2, T = 500


def test_pocket_sharding_covers_every_pocket_once():
    from diffndm_b200.parallel import shard_pockets
    for n, w in ((100, 8), (7, 3), (3, 8), (0, 4)):
        got = sorted(i for r in range(w) for i in shard_pockets(n, r, w))
        assert got == list(range(n))
