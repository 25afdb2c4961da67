"""GPU: the engine on the round-2 fixtures generated from the reference (tests/golden/make_golden_r2.py).

Bars (BASELINE north_star), asserted in absolute units because the score-wrapped trajectories keep |z| < 10 A: per-step
coordinates within 1e-3 A, features within 1e-2 relative; eps_x within max(1e-3, 2e-3 |eps_x|max)."""
import json
import os

import numpy as np
import pytest
import torch

from r2_common import R2, FWD, TRAJ, INP, weights_for, traj_draws, inpaint_draws, batch_of
from guidance_common import SyntheticScoreDynamics

pytestmark = pytest.mark.gpu
STEP_X_TOL = 1e-3        # A, absolute
H_REL_TOL = 1e-2
REPORT = {}


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda', 0)


def _engine(W, **kw):
    from diffndm_b200.engine import B200EGNNDynamics
    from diffndm_b200.weights import DynamicsConfig
    return B200EGNNDynamics(DynamicsConfig(), W, max_nodes=4096, max_edges=300000, max_samples=64, **kw).eval()


@pytest.fixture(scope='module')
def dyn(golden_weights, dev):
    return _engine(golden_weights)


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize('name', FWD)
def test_forward_vs_reference_f64(name, golden_weights, dyn, dev):
    c = R2[name]
    W = weights_for(name, golden_weights)
    d = dyn if W is golden_weights else _engine(W)
    N, n_l = len(c['lig_mask']) + len(c['pocket_mask']), len(c['lig_mask'])
    trh, _ = d.engine.set_trace(N)
    out_l, _ = d(_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['t'], dev), _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev))
    out = out_l.cpu().numpy()
    ref = c['out_lig_f64']
    ex, sx = np.abs(out[:, :3] - ref[:, :3]).max(), np.abs(ref[:, :3]).max()
    eh, sh = np.abs(out[:, 3:] - ref[:, 3:]).max(), np.abs(ref[:, 3:]).max()
    h = trh[-1, :n_l].cpu().numpy()
    eh6 = np.abs(h - c['h_lig_last']).max() / np.abs(c['h_lig_last']).max()
    d.engine.clear_trace()
    REPORT[name] = dict(eps_x_err=float(ex), eps_x_scale=float(sx), eps_h_rel=float(eh / sh), h_last_rel=float(eh6))
    assert ex < max(1e-3, 2e-3 * sx), f'eps_x err {ex:.3e} (scale {sx:.3f})'
    assert eh < H_REL_TOL * sh, f'eps_h err {eh:.3e} (scale {sh:.3f})'
    assert eh6 < H_REL_TOL, f'h after the last block: {eh6:.3e} relative'


def _sampler(dyn, c, dev):
    from diffndm_b200.sampler import ConditionalSampler
    return ConditionalSampler(SyntheticScoreDynamics(dyn, c['x0_rel']).to(dev), timesteps=500)


@pytest.mark.parametrize('name', TRAJ)
def test_teacher_forced_steps_absolute(name, dyn, dev):
    """sample_p_zs_given_zt (conditional_model.py:483-540) from the reference's recorded states with its draws."""
    c = R2[name]
    B, n_p, lm, pm = batch_of(c)
    Tn = int(c['timesteps'])
    smp = _sampler(dyn, c, dev)
    draws = traj_draws(c)
    worst = 0.0
    for s in [int(v) for v in c['kept_steps']]:
        s_arr = torch.full((B, 1), s, dtype=torch.float32) / Tn
        t_arr = torch.full((B, 1), s + 1, dtype=torch.float32) / Tn
        z, xp = smp.sample_p_zs_given_zt(s_arr, t_arr, _t(c[f's{s}_z_in'], dev), _t(c[f's{s}_xp_in'], dev), _t(lm, dev), _t(pm, dev),
                                         noise=_t(draws[Tn - s], dev), n_samples=B)
        z, xp = z.cpu().numpy(), xp.cpu().numpy()
        rz, rp = c[f's{s}_z_out'], c[f's{s}_xp_out']
        dx = np.abs(z[:, :3] - rz[:, :3]).max()
        dp = np.abs(xp[:, :3] - rp[:, :3]).max()
        dh = np.abs(z[:, 3:] - rz[:, 3:]).max() / max(1.0, np.abs(rz[:, 3:]).max())
        worst = max(worst, dx, dp)
        assert dx < STEP_X_TOL and dp < STEP_X_TOL, f's={s}: coordinates off by {dx:.2e} A (pocket {dp:.2e})'
        assert dh < H_REL_TOL, f's={s}: features off by {dh:.2e}'
    REPORT[name + '/teacher_forced_max_dx'] = float(worst)


@pytest.mark.parametrize('name', TRAJ)
def test_free_running_trajectory_absolute(name, dyn, dev):
    """The whole sample_given_pocket on the reference's draws, its own states all the way: the synthetic score contracts
    errors like a trained denoiser does, so the end state stays within a few 1e-3 A of the reference's."""
    c = R2[name]
    B, n_p, lm, pm = batch_of(c)
    Tn = int(c['timesteps'])
    smp = _sampler(dyn, c, dev)
    oh = np.eye(10, dtype=np.float32)[c['pocket_t']]
    pocket = {'x': torch.from_numpy(np.tile(c['pocket_x'], (B, 1))), 'one_hot': torch.from_numpy(np.tile(oh, (B, 1))),
              'size': torch.tensor([n_p] * B), 'mask': torch.from_numpy(pm)}
    noise = _t(np.stack(traj_draws(c)), dev)
    xh_l, xh_p, _, _ = smp.sample_given_pocket(pocket, c['sizes'], timesteps=Tn, noise=noise)
    ref = c['final_lig']
    dx = np.abs(xh_l.cpu().numpy()[:, :3] - ref[:, :3]).max()
    REPORT[name + '/free_running_final_dx'] = float(dx)
    assert dx < 5e-3, dx
    assert np.array_equal(xh_l.cpu().numpy()[:, 3:].argmax(1), ref[:, 3:].argmax(1))
    assert np.abs(xh_p.cpu().numpy()[:, :3] - c['final_pocket'][:, :3]).max() < 5e-3


@pytest.mark.parametrize('name', INP)
def test_inpaint_large_pockets_absolute(name, dyn, dev):
    """ConditionalSampler.inpaint on the 5ndu pocket (fragments fixed) and a 600-atom pocket (BASELINE configs[4])."""
    c = R2[name]
    B, n_p, _, pm = batch_of(c)
    smp = _sampler(dyn, c, dev)
    oh = np.eye(10, dtype=np.float32)
    pocket = {'x': torch.from_numpy(np.tile(c['pocket_x'], (B, 1))), 'one_hot': torch.from_numpy(np.tile(oh[c['pocket_t']], (B, 1))),
              'size': torch.tensor([n_p] * B), 'mask': torch.from_numpy(pm)}
    ligand = {'x': torch.from_numpy(c['lig_x']), 'one_hot': torch.from_numpy(oh[c['lig_t']]),
              'size': torch.from_numpy(c['sizes']), 'mask': torch.from_numpy(c['lig_mask'])}
    xh_l, xh_p, lm, _ = smp.inpaint(ligand, pocket, torch.from_numpy(c['lig_fixed']), resamplings=int(c['resamplings']),
                                    timesteps=int(c['timesteps']), noise=[_t(n, dev) for n in inpaint_draws(c)])
    ref = c['final_lig']
    dx = np.abs(xh_l.cpu().numpy()[:, :3] - ref[:, :3]).max()
    REPORT[name + '/final_dx'] = float(dx)
    assert dx < 5e-3, dx
    assert (xh_l.cpu().numpy()[:, 3:].argmax(1) == ref[:, 3:].argmax(1)).mean() >= 0.98
    assert np.abs(xh_p.cpu().numpy()[:, :3] - c['final_pocket'][:, :3]).max() < 5e-3
    assert np.array_equal(lm.cpu().numpy(), c['lig_mask'])


def test_zz_write_report():
    """Measured errors of this run -> gpurun_out/r2_parity_errors[_f32radial].json (copied into profiles/ by hand)."""
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    tag = '_f32radial' if os.environ.get('DNDM_GCL_F32_RADIAL') == '1' else ''
    with open(os.path.join(out, f'r2_parity_errors{tag}.json'), 'w') as f:
        json.dump(REPORT, f, indent=1)
