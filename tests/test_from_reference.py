"""The drop-in constructors.  CPU: ``B200EGNNDynamics.from_reference`` reads the REAL reference module (build container
only, skipped where /root/reference is absent) and the structural stand-in identically.  GPU: the stand-in goes through
``from_reference`` / ``B200ConditionalDDPM.from_reference`` and the reference-shaped calls reproduce the golden vectors."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_npz_groups

sys.path.insert(0, GOLDEN)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_standin  # noqa: E402
from ref_loader import reference_available  # noqa: E402


class _FakeEngine:
    """Stands in for engine.Engine on a box without a GPU: records what from_reference hands it."""
    last = None

    def __init__(self, cfg, max_nodes, max_edges, max_samples):
        self.cfg, self.state, self.device = cfg, None, 0
        _FakeEngine.last = self

    def load_weights(self, state):
        self.state = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in state.items()}


def _from_reference_cpu(module, monkeypatch):
    from diffndm_b200 import engine as E
    monkeypatch.setattr(E, 'Engine', _FakeEngine)
    d = E.B200EGNNDynamics.from_reference(module)
    return d.cfg, _FakeEngine.last.state


def test_standin_matches_expected_keys(golden_weights, monkeypatch):
    from diffndm_b200.weights import DynamicsConfig, expected_keys
    cfg = DynamicsConfig()
    m = ref_standin.build(cfg, golden_weights)
    got_cfg, state = _from_reference_cpu(m, monkeypatch)
    assert got_cfg == cfg
    for name, shape in expected_keys(cfg):
        assert tuple(state[name].shape) == tuple(shape), name
        assert np.array_equal(state[name], golden_weights[name]), name


@pytest.mark.skipif(not reference_available(), reason='/root/reference is only present in the build container')
def test_from_reference_on_the_real_reference_module(golden_weights, monkeypatch):
    from ref_loader import build_reference_model
    from diffndm_b200.weights import DynamicsConfig, expected_keys
    cfg = DynamicsConfig()
    dyn, ddpm = build_reference_model(cfg, golden_weights)
    got_cfg, state = _from_reference_cpu(dyn, monkeypatch)
    assert got_cfg == cfg                                     # every hyper-parameter recovered from the live module
    for name, shape in expected_keys(cfg):
        assert np.array_equal(state[name], golden_weights[name]), name
    # the stand-in used on the GPU box exposes the same state-dict keys and the attributes from_reference reads
    sm = ref_standin.build(cfg, golden_weights)
    assert set(sm.state_dict().keys()) == set(dyn.state_dict().keys())
    for attr in ('n_dims', 'edge_cutoff_l', 'edge_cutoff_p', 'edge_cutoff_i', 'edge_nf', 'update_pocket_coords', 'condition_time'):
        assert getattr(sm, attr) == getattr(dyn, attr), attr
    # ... and the schedule table the adapter takes over is the one the sampler computes itself
    from diffndm_b200.sampler import polynomial_gamma
    assert torch.equal(ddpm.gamma.gamma.detach().cpu(), polynomial_gamma(500, 5.0e-4, 2.0))
    assert int(ddpm.T) == 500 and list(ddpm.norm_values) == [1, 4]


@pytest.mark.skipif(not reference_available(), reason='/root/reference is only present in the build container')
@pytest.mark.parametrize('name', ['moad192', 'narrow128', 'ca20'])
def test_from_reference_recovers_the_other_configurations(name, monkeypatch):
    """hidden_nf 192 + edge-type embedding + 4 / 7 A cutoffs (moad_fullatom_cond), 128 / joint 32 / five blocks, residue_nf 20:
    every hyper-parameter is read back from the live reference module, the state dict (``edge_embedding.weight`` included)
    is handed to the engine unchanged, and ``engine_table`` accepts it."""
    from ref_loader import build_reference_model
    from make_golden_widen import case_config
    from diffndm_b200.weights import engine_table, expected_keys, random_init
    cfg = case_config(name)
    W = random_init(cfg, 7, 0.3)
    dyn, _ = build_reference_model(cfg, W)
    got_cfg, state = _from_reference_cpu(dyn, monkeypatch)
    assert got_cfg == cfg
    for key, shape in expected_keys(cfg):
        assert tuple(state[key].shape) == tuple(shape) and np.array_equal(state[key], W[key]), key
    ecfg, tab = engine_table(got_cfg, state)
    assert ecfg.hidden_nf == 256 and (('egnn.e_block_0.gcl_0.edge_mlp.0.edge_type_bias' in tab) == bool(cfg.edge_embedding_dim))


# ------------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------------
FWD_CASES, _ = load_npz_groups('forward.npz')
TRAJ_CASES, _ = load_npz_groups('trajectory.npz')


@pytest.fixture(scope='module')
def adapter(golden_weights):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from diffndm_b200.adapter import B200ConditionalDDPM
    from diffndm_b200.sampler import polynomial_gamma
    from diffndm_b200.weights import DynamicsConfig
    ddpm = ref_standin.StandInDDPM(ref_standin.build(DynamicsConfig(), golden_weights), polynomial_gamma(500, 5.0e-4, 2.0))
    return B200ConditionalDDPM.from_reference(ddpm, max_nodes=4096, max_edges=300000, max_samples=64)


@pytest.mark.gpu
def test_from_reference_forward_matches_golden(adapter):
    c = FWD_CASES['3rfm_b2']
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(a).to(dev)
    out_l, out_p = adapter.dynamics(t(c['xh_lig']), t(c['xh_pocket']), t(c['t']), t(c['lig_mask']), t(c['pocket_mask']))
    ref = c['out_lig_f64']
    assert np.abs(out_l.cpu().numpy()[:, :3] - ref[:, :3]).max() < max(1e-3, 2e-3 * np.abs(ref[:, :3]).max())
    assert np.abs(out_l.cpu().numpy()[:, 3:] - ref[:, 3:]).max() < 1e-2 * np.abs(ref[:, 3:]).max()


@pytest.mark.gpu
def test_adapter_step_returns_reference_triple(adapter, monkeypatch):
    """sample_p_zs_given_zt(s, t, zt_lig, xh0_pocket, ligand_mask, pocket_mask, optimize) -> (zs, xh_pocket, log_prob_adjust)
    (conditional_model.py:483-540), on the reference's recorded state with its draw."""
    c = TRAJ_CASES['3rfm_b2_T5']
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(a).to(dev)
    noise = t(c['noise'][1])
    monkeypatch.setattr(torch, 'randn', lambda *a, **k: noise.clone())
    out = adapter.sample_p_zs_given_zt(torch.from_numpy(c['step0/s']), torch.from_numpy(c['step0/t']), t(c['step0/z_in']),
                                       t(c['step0/xp_in']), t(c['lig_mask']), t(c['pocket_mask']), 0)
    assert len(out) == 3 and out[2].numel() == 1
    scale = max(1.0, np.abs(c['step0/z_out'][:, :3]).max())
    assert np.abs(out[0].cpu().numpy() - c['step0/z_out'])[:, :3].max() / scale < 1e-3
    assert np.abs(out[1].cpu().numpy() - c['step0/xp_out'])[:, :3].max() / scale < 1e-3
    with pytest.raises(NotImplementedError):
        adapter.sample_p_zs_given_zt(torch.from_numpy(c['step0/s']), torch.from_numpy(c['step0/t']), t(c['step0/z_in']),
                                     t(c['step0/xp_in']), t(c['lig_mask']), t(c['pocket_mask']), 1)


@pytest.mark.gpu
def test_adapter_takes_the_reference_positional_arguments(adapter):
    """lightning_modules.py:899-901 / conditional_model.py:886-887: 14 positionals + timesteps; same result as the
    sampler's own entry point on the same seed."""
    from diffndm_b200 import synthetic
    from diffndm_b200.hostpool import PooledReward, radius_of_gyration_score
    px, pt = synthetic.synthetic_pocket(3, 40)
    B = 3
    oh = np.eye(10, dtype=np.float32)[pt]
    mk = lambda: {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(oh, (B, 1))),
                  'size': torch.tensor([len(px)] * B), 'mask': torch.arange(B).repeat_interleave(len(px))}
    sizes = torch.tensor([5, 8, 6])
    com = torch.zeros(B, 3)
    adapter.sampler.sample_given_pocket(mk(), sizes, timesteps=20)          # captures the reverse-step graph for this batch shape
    torch.manual_seed(5)
    a = adapter.sample_given_pocket(mk(), sizes, com, None, False, 0, False, 'x', 'cuda', 0, None, None, 0, 0, timesteps=20)
    torch.manual_seed(5)
    b = adapter.sampler.sample_given_pocket(mk(), sizes, timesteps=20)
    assert len(a) == 4 and all(torch.equal(x, y) for x, y in zip(a, b))
    assert a[0].shape == (19, 13) and a[1].shape == (B * len(px), 13)
    # guided modes through the same signature with a plugged-in scorer (RDKit is absent on this image)
    adapter.reward_fn = PooledReward(radius_of_gyration_score, workers=0)
    g = adapter.sample_given_pocket(mk(), sizes, com, None, False, 0, False, 'x', 'cuda', 0, None, None, 1, 1, timesteps=60)
    adapter.reward_fn = None
    assert g[0].shape[1] == 13 and torch.isfinite(g[0][:, :3]).all()
    with pytest.raises(NotImplementedError):
        adapter.sample_given_pocket(mk(), sizes, com, None, False, 0, False, 'x', 'cuda', 1, None, None, 0, 0, timesteps=20)
