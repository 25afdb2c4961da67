"""CPU: the numpy oracle replays every golden vector generated from the reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import load_npz_groups
from oracle import egnn_oracle as O

CFG = O.OracleConfig()
EDGE_CASES, _ = load_npz_groups('edges.npz')
FWD_CASES, _ = load_npz_groups('forward.npz')
TRAJ_CASES, _ = load_npz_groups('trajectory.npz')


@pytest.mark.parametrize('name', sorted(EDGE_CASES))
def test_edges_bit_exact(name):
    c = EDGE_CASES[name]
    e = O.get_edges(c['lig_mask'], c['pocket_mask'], c['xh_lig'][:, :3], c['xh_pocket'][:, :3], CFG)
    ref = c['edges'].astype(np.int64)
    # bit-exact set AND order (dynamics.py:183-185).  n_boundary pairs lie within 1e-4 A of the cutoff in
    # fp64; the reference's cdist and the oracle's direct differences agreed on all of them at generation time.
    assert e.shape == ref.shape, (e.shape, ref.shape, int(c['n_boundary']))
    assert np.array_equal(e, ref)
    # receiver-sorted, cols ascending, self loops present
    assert np.all(np.diff(e[0]) >= 0)
    n = len(c['lig_mask']) + len(c['pocket_mask'])
    assert np.array_equal(np.unique(e[0][e[0] == e[1]]), np.arange(n))


@pytest.mark.parametrize('name', sorted(FWD_CASES))
def test_forward_matches_reference(name, golden_weights):
    c = FWD_CASES[name]
    trace = {}
    ol, op = O.dynamics_forward(golden_weights, c['xh_lig'], c['xh_pocket'], c['t'], c['lig_mask'],
                                c['pocket_mask'], CFG, dtype=np.float32, trace=trace)
    # fp32 restatement vs fp32 reference: only summation order differs
    assert np.abs(ol - c['out_lig_f32']).max() < 2e-5
    assert np.abs(op - c['out_pocket_f32']).max() < 2e-5
    assert np.abs(op[:, :3]).max() == 0.0          # pocket velocity is exactly zero
    o64l, o64p = O.dynamics_forward(golden_weights, c['xh_lig'], c['xh_pocket'], c['t'], c['lig_mask'],
                                    c['pocket_mask'], CFG, dtype=np.float64, trace=trace)
    assert np.abs(o64l - c['out_lig_f64']).max() < 1e-9
    assert np.abs(o64p - c['out_pocket_f64']).max() < 1e-9
    rows = c['trace_rows']
    n_l = len(c['lig_mask'])
    for i in range(CFG.n_layers):
        assert np.abs(trace[f'h_{i}'][rows] - c[f'h_rows_{i}']).max() < 1e-9
        assert np.abs(trace[f'x_{i}'][:n_l] - c[f'x_lig_{i}']).max() < 1e-9


def test_gamma_table_properties():
    g = O.gamma_table(500, 5e-4, 2.0)
    assert g.shape == (501,) and g.dtype == np.float32
    assert np.all(np.diff(g) > 0)
    sc = O.step_scalars(g[:-1], g[1:])
    assert np.all(sc['sigma2_ts'] > 0) and np.all(sc['alpha_ts'] <= 1.0)


@pytest.mark.parametrize('name', sorted(TRAJ_CASES))
def test_sampler_steps_match_reference(name, golden_weights):
    c = TRAJ_CASES[name]
    Tn = int(c['timesteps'])
    g = O.gamma_table(500, 5e-4, 2.0)
    lm, pm = c['lig_mask'], c['pocket_mask']
    for i in range(Tn):
        s, t = c[f'step{i}/s'][:, 0], c[f'step{i}/t'][:, 0]
        gs = g[np.round(s * 500).astype(np.int64)]
        gt = g[np.round(t * 500).astype(np.int64)]
        eps, _ = O.dynamics_forward(golden_weights, c[f'step{i}/z_in'], c[f'step{i}/xp_in'], t[:, None], lm, pm, CFG)
        z, xp = O.sample_p_zs_given_zt(c[f'step{i}/z_in'], c[f'step{i}/xp_in'], eps, c['noise'][1 + i], gs, gt, lm, pm)
        scale = max(1.0, np.abs(c[f'step{i}/z_out']).max())
        assert np.abs(z - c[f'step{i}/z_out']).max() < 2e-5 * scale
        assert np.abs(xp - c[f'step{i}/xp_out']).max() < 2e-5 * scale
    # final p(x,h|z0) head (conditional_model.py:1427, 136-160)
    z0, xp0 = c[f'step{Tn - 1}/z_out'], c[f'step{Tn - 1}/xp_out']
    B = len(c['sizes'])
    eps0, _ = O.dynamics_forward(golden_weights, z0, xp0, np.zeros((B, 1), np.float32), lm, pm, CFG)
    x, types, xpk, hpk = O.sample_p_xh_given_z0(z0, xp0, eps0, c['noise'][Tn + 1], np.full(B, g[0]), lm, pm, CFG)
    assert np.array_equal(types, c['final_lig'][:, 3:].argmax(1))
    scale = max(1.0, np.abs(c['final_lig'][:, :3]).max())
    # the reference may re-project on CoG drift (:1432-1438); compare up to that projection
    ref_x = c['final_lig'][:, :3]
    assert np.abs(x - ref_x).max() < 5e-4 * scale


INPAINT_CASES, _ = load_npz_groups('inpaint.npz')


def _inpaint_inputs(c):
    B, n_p = len(c['sizes']), len(c['pocket_x'])
    oh = np.eye(10, dtype=np.float32)
    return dict(lig_x=c['lig_x'], lig_onehot=oh[c['lig_t']], lig_mask=c['lig_mask'], pocket_x=np.tile(c['pocket_x'], (B, 1)),
                pocket_onehot=np.tile(oh[c['pocket_t']], (B, 1)), pocket_mask=np.repeat(np.arange(B), n_p),
                lig_fixed=c['lig_fixed'])


@pytest.mark.parametrize('name', sorted(INPAINT_CASES))
def test_inpaint_matches_reference(name, golden_weights):
    """Free-running RePaint loop (conditional_model.py:1491-1790) with the reference's Gaussian draws injected."""
    c = INPAINT_CASES[name]
    a = _inpaint_inputs(c)
    xh_l, xh_p, z0, xp0 = O.inpaint(golden_weights, a['lig_x'], a['lig_onehot'], a['lig_mask'], a['pocket_x'],
                                    a['pocket_onehot'], a['pocket_mask'], a['lig_fixed'], c['noise'], int(c['timesteps']),
                                    int(c['resamplings']), CFG)
    scale = max(1.0, np.abs(c['z_final_in']).max())
    assert np.abs(z0 - c['z_final_in']).max() < 2e-5 * scale
    assert np.abs(xp0 - c['xp_final_in']).max() < 2e-5 * scale
    assert np.array_equal(xh_l[:, 3:].argmax(1), c['final_lig'][:, 3:].argmax(1))
    assert np.abs(xh_l[:, :3] - c['final_lig'][:, :3]).max() < 2e-5 * scale
    assert np.abs(xh_p - c['final_pocket']).max() < 2e-5 * scale


def test_inpaint_moves_are_com_free_and_keep_known_atoms():
    """q(z_s|x) and the re-noising move project onto the ligand-COM-free subspace; with sigma -> 0 the blend returns the
    (translated) input on the fixed atoms."""
    rng = np.random.default_rng(5)
    lm = np.repeat(np.arange(3), [4, 7, 5])
    pm = np.repeat(np.arange(3), 6)
    xh_l = rng.standard_normal((16, 13)).astype(np.float32)
    xh_p = rng.standard_normal((18, 13)).astype(np.float32)
    g = O.gamma_table()
    z, xp = O.noised_representation(xh_l, xh_p, rng.standard_normal((16, 13)), g[[5, 100, 400]], lm, pm)
    assert np.abs(O.segment_mean(z[:, :3], lm, 3)).max() < 1e-5
    z2, xp2 = O.sample_p_zt_given_zs(z, xp, rng.standard_normal((16, 13)), g[[6, 101, 401]], g[[5, 100, 400]], lm, pm)
    assert np.abs(O.segment_mean(z2[:, :3], lm, 3)).max() < 1e-5
    # pocket and ligand are shifted by the same per-sample vector
    d_l = O.segment_mean(z[:, :3] - np.sqrt(O.sigmoid(-g[[5, 100, 400]]))[lm][:, None] * xh_l[:, :3], lm, 3)
    fixed = (rng.random(16) < 0.5).astype(np.float32)
    fixed[[0, 4, 11]] = 1
    zc, xpc = O.inpaint_combine(z, z2, xp, fixed, lm, pm, 3)
    assert np.array_equal(zc[fixed == 0], z2[fixed == 0])
    fx = fixed > 0
    assert np.abs(O.segment_mean(zc[fx][:, :3], lm[fx], 3) - O.segment_mean(z2[fx][:, :3], lm[fx], 3)).max() < 1e-5
    assert d_l.shape == (3, 3)


BOND_CASES, BOND_META = load_npz_groups('bonds.npz')


@pytest.mark.parametrize('name', sorted(BOND_CASES))
def test_bond_orders_match_reference(name):
    """get_bond_order_batch as called by make_mol_edm (analysis/molecule_builder.py:30-55, 100-113): bit-exact."""
    c = BOND_CASES[name]
    assert int(c['pairs_near_threshold']) == 0          # no pair where cdist and direct differences could disagree
    mats, val, stats = O.bond_orders(c['x'], c['types'], c['mask'], BOND_META['bonds1'], BOND_META['bonds2'],
                                     BOND_META['bonds3'], BOND_META['margins'])
    assert np.array_equal(np.concatenate([m.reshape(-1) for m in mats]), c['e_flat'])
    # derived quantities are consistent with E
    off = 0
    for b, n in enumerate(c['sizes']):
        e = c['e_flat'][off:off + n * n].reshape(n, n).astype(np.int32)
        off += n * n
        assert np.array_equal(val[c['mask'] == b], (e + e.T).sum(1))
        assert stats[b, 0] == int((e > 0).sum()) and 1 <= stats[b, 1] <= n and stats[b, 2] <= n
