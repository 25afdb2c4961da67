"""CPU: the numpy oracle on the round-2 fixtures generated from the reference (tests/golden/make_golden_r2.py): untied
coordinate heads, the radial stress case, the 5ndu and 600-atom pockets, score-wrapped trajectories and inpainting."""
import numpy as np
import pytest

from r2_common import R2, FWD, TRAJ, INP, weights_for, traj_draws, inpaint_draws, batch_of
from guidance_common import polynomial2_gamma, score_numpy
from oracle import egnn_oracle as O

CFG = O.OracleConfig()
GAM = polynomial2_gamma().numpy()


@pytest.mark.parametrize('name', FWD)
def test_forward_f64_matches_reference(name, golden_weights):
    c = R2[name]
    W = weights_for(name, golden_weights)
    trace = {}
    ol, _ = O.dynamics_forward(W, c['xh_lig'], c['xh_pocket'], c['t'], c['lig_mask'], c['pocket_mask'], CFG, dtype=np.float64,
                               trace=trace)
    assert np.abs(ol - c['out_lig_f64']).max() < 1e-8 * max(1.0, np.abs(c['out_lig_f64']).max())
    n_l = len(c['lig_mask'])
    assert np.abs(trace[f'h_{CFG.n_layers - 1}'][:n_l] - c['h_lig_last']).max() < 1e-5 * np.abs(c['h_lig_last']).max()


def _eps_fn(W, c):
    def f(z, xp, t, lm, pm):
        eps, _ = O.dynamics_forward(W, z, xp, np.asarray(t, np.float32), lm, pm, CFG)
        return score_numpy(eps.astype(np.float32), z, xp, t, lm, pm, c['x0_rel'], GAM)
    return f


@pytest.mark.parametrize('name', TRAJ)
def test_score_wrapped_steps_match_reference(name, golden_weights):
    c = R2[name]
    B, n_p, lm, pm = batch_of(c)
    Tn = int(c['timesteps'])
    draws = traj_draws(c)
    f = _eps_fn(golden_weights, c)
    g = O.gamma_table(500)
    for s in [int(v) for v in c['kept_steps']][::3]:
        t_s = np.full((B, 1), s, np.float32) / np.float32(Tn)
        t_t = (np.full((B, 1), s, np.float32) + np.float32(1)) / np.float32(Tn)
        assert int(c[f's{s}_d0']) == Tn - s
        eps = f(c[f's{s}_z_in'], c[f's{s}_xp_in'], t_t, lm, pm)
        look = lambda t: g[np.round(t.reshape(-1) * 500).astype(np.int64)]
        z, xp = O.sample_p_zs_given_zt(c[f's{s}_z_in'], c[f's{s}_xp_in'], eps, draws[Tn - s], look(t_s), look(t_t), lm, pm)
        assert np.abs(z - c[f's{s}_z_out']).max() < 3e-5 and np.abs(xp - c[f's{s}_xp_out']).max() < 3e-5, s
        assert np.abs(c[f's{s}_z_in'][:, :3]).max() < 10.0                  # realistic states: absolute tolerances mean something


@pytest.mark.parametrize('name', INP[:1])
def test_score_wrapped_inpaint_matches_reference(name, golden_weights):
    c = R2[name]
    B, n_p, _, pm = batch_of(c)
    oh = np.eye(10, dtype=np.float32)
    xh_l, xh_p, _, _ = O.inpaint(golden_weights, c['lig_x'], oh[c['lig_t']], c['lig_mask'], np.tile(c['pocket_x'], (B, 1)),
                                 np.tile(oh[c['pocket_t']], (B, 1)), pm, c['lig_fixed'], inpaint_draws(c), int(c['timesteps']),
                                 int(c['resamplings']), CFG, eps_fn=_eps_fn(golden_weights, c))
    assert np.abs(xh_l[:, :3] - c['final_lig'][:, :3]).max() < 1e-4
    assert np.array_equal(xh_l[:, 3:].argmax(1), c['final_lig'][:, 3:].argmax(1))
    assert np.abs(xh_p[:, :3] - c['final_pocket'][:, :3]).max() < 1e-4
