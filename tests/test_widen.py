"""The other pocket-conditional configurations of the reference (tests/golden/widen.npz, generated from the unmodified
reference by tests/golden/make_golden_widen.py): hidden_nf 192 with the edge-type embedding and 4 A / 7 A cutoffs
(moad_fullatom_cond), hidden_nf 128 / joint_nf 32 / five blocks, and residue_nf 20 != atom_nf (C-alpha pockets).

CPU: the numpy oracle against the reference's fp64 output, and the exactness of ``weights.engine_table`` (zero-padding to the
compiled width + folding the embedding: the oracle on the REWRITTEN table must reproduce the oracle on the original one).
GPU: the engine against the reference's edges (bit-exact) and fp64 output."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden_widen import CASES, case_config  # noqa: E402

from diffndm_b200.weights import engine_table, random_init, weights_checksum  # noqa: E402
from oracle import egnn_oracle as O  # noqa: E402

FX = np.load(os.path.join(GOLDEN, 'widen.npz'))


def _case(name):
    c = {k.split('/', 1)[1]: FX[k] for k in FX.files if k.startswith(name + '/')}
    cfg = case_config(name)
    W = random_init(cfg, 7, 0.3)
    assert abs(weights_checksum(W) - float(c['weights_checksum'])) < 1e-6 * max(1.0, abs(float(c['weights_checksum'])))
    return c, cfg, W


def _ocfg(cfg):
    return O.OracleConfig(atom_nf=cfg.atom_nf, residue_nf=cfg.residue_nf, joint_nf=cfg.joint_nf, hidden_nf=cfg.hidden_nf,
                          n_layers=cfg.n_layers, edge_cutoff_ligand=cfg.edge_cutoff_ligand, edge_cutoff_pocket=cfg.edge_cutoff_pocket,
                          edge_cutoff_interaction=cfg.edge_cutoff_interaction)


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_matches_reference(name):
    c, cfg, W = _case(name)
    oc = _ocfg(cfg)
    e = O.get_edges(c['lig_mask'], c['pocket_mask'], c['xh_lig'][:, :3], c['xh_pocket'][:, :3], oc)
    assert np.array_equal(e, c['edges'].astype(np.int64))
    out_l, out_p = O.dynamics_forward(W, c['xh_lig'], c['xh_pocket'], c['t'], c['lig_mask'], c['pocket_mask'], oc, dtype=np.float64)
    assert np.abs(out_l - c['out_lig_f64']).max() < 1e-9
    assert np.abs(out_p - c['out_pocket_f64']).max() < 1e-9


@pytest.mark.parametrize('name', sorted(CASES))
def test_engine_table_is_exact(name):
    """Zero-padding to 256 channels and folding the edge-type embedding into per-type hidden vectors do not change the
    function: the oracle evaluated on the rewritten table (with the type vectors added to the first-layer pre-activation,
    as the edge kernel does) equals the oracle on the reference's table to fp64 round-off."""
    c, cfg, W = _case(name)
    ecfg, tab = engine_table(cfg, W)
    assert ecfg.hidden_nf == 256 and ecfg.edge_embedding_dim is None
    ref_l, _ = O.dynamics_forward(W, c['xh_lig'], c['xh_pocket'], c['t'], c['lig_mask'], c['pocket_mask'], _ocfg(cfg), dtype=np.float64)
    # the rewritten table in the oracle's vocabulary: per-type vectors = an embedding of width 3 (one-hot) whose first-layer
    # columns are the vectors themselves
    W2 = {k: v for k, v in tab.items() if not k.endswith('edge_type_bias')}
    if cfg.edge_embedding_dim:
        W2['edge_embedding.weight'] = np.eye(3, dtype=np.float32)
        for k, v in tab.items():
            if k.endswith('edge_type_bias'):
                wk = k.replace('edge_type_bias', 'weight')
                W2[wk] = np.concatenate([W2[wk], v.T], axis=1)
    got_l, _ = O.dynamics_forward(W2, c['xh_lig'], c['xh_pocket'], c['t'], c['lig_mask'], c['pocket_mask'], _ocfg(ecfg), dtype=np.float64)
    assert np.abs(got_l - ref_l).max() < 1e-7 * max(1.0, np.abs(ref_l).max())
    # the padding is inert: padded hidden rows of every packed matrix are exactly zero
    H = cfg.hidden_nf
    assert not np.any(tab['egnn.e_block_0.gcl_0.edge_mlp.2.weight'][H:]) and not np.any(tab['egnn.e_block_0.gcl_0.edge_mlp.2.weight'][:, H:])


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(CASES))
def test_engine_matches_reference(name):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from diffndm_b200.engine import B200EGNNDynamics
    dev = torch.device('cuda', 0)
    c, cfg, W = _case(name)
    dyn = B200EGNNDynamics(cfg, W, max_nodes=2048, max_edges=65536, max_samples=16).eval()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    B = len(c['t'])
    rp, col = dyn.engine.radius_graph(t(c['xh_lig']), t(c['xh_pocket']), t(c['lig_mask']), t(c['pocket_mask']), B)
    rp, col = rp.cpu().numpy(), col.cpu().numpy()
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    assert np.array_equal(np.stack([rows, col.astype(np.int64)]), c['edges'].astype(np.int64))      # cutoffs 4 / 7 A included
    out_l, out_p = dyn(t(c['xh_lig']), t(c['xh_pocket']), t(c['t']), t(c['lig_mask']), t(c['pocket_mask']), n_samples=B)
    out_l, out_p = out_l.cpu().numpy(), out_p.cpu().numpy()
    ref = c['out_lig_f64']
    ex, sx = np.abs(out_l[:, :3] - ref[:, :3]).max(), np.abs(ref[:, :3]).max()
    # coordinates: the north_star bar is 1e-3 A per sampling step; a step scales the denoiser output by c_eps <= 0.177
    # (polynomial_2, T = 500), so eps_x must be within 5.6e-3.  Measured 1.5e-3 at |eps_x| = 0.56 on moad192 (48 pocket
    # neighbours per ligand atom at the 7 A interaction cutoff: more bf16-rounded terms per coordinate sum than at 5 A)
    assert ex * 0.177 < 1e-3 and ex < max(1e-3, 4e-3 * sx), (ex, sx)
    assert np.abs(out_l[:, 3:] - ref[:, 3:]).max() < 1e-2 * max(1.0, np.abs(ref[:, 3:]).max())     # features, 1e-2 relative
    assert np.abs(out_p[:, 3:] - c['out_pocket_f64'][:, 3:]).max() < 1e-2 * max(1.0, np.abs(c['out_pocket_f64'][:, 3:]).max())
    assert np.abs(out_p[:, :3]).max() == 0.0                                                       # the pocket is frozen


@pytest.mark.gpu
def test_sampler_step_with_wider_pocket_rows():
    """C-alpha pockets: pocket rows are [3 + 20] wide, ligand rows [3 + 10].  One reverse step through the fused kernel against
    the oracle's p(z_s | z_t) on the engine's own eps, then a short trajectory through the public sampler."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from diffndm_b200.engine import B200EGNNDynamics
    from diffndm_b200.sampler import ConditionalSampler
    dev = torch.device('cuda', 0)
    c, cfg, W = _case('ca20')
    dyn = B200EGNNDynamics(cfg, W, max_nodes=2048, max_edges=65536, max_samples=16).eval()
    smp = ConditionalSampler(dyn, timesteps=500)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    B = len(c['t'])
    lm, pm = t(c['lig_mask']), t(c['pocket_mask'])
    # a COM-free state: the sampler kernel projects ligand and pocket by the ligand's centre of mass
    z = c['xh_lig'].copy()
    xp = c['xh_pocket'].copy()
    for b in range(B):
        com = z[c['lig_mask'] == b, :3].mean(0)
        z[c['lig_mask'] == b, :3] -= com
        xp[c['pocket_mask'] == b, :3] -= com
    s_arr = np.full((B, 1), 249 / 500, np.float32)
    t_arr = np.full((B, 1), 250 / 500, np.float32)
    rng = np.random.default_rng(5)
    noise = rng.standard_normal(z.shape).astype(np.float32)
    eps, _ = dyn(t(z), t(xp), t(t_arr), lm, pm, n_samples=B)
    zs, xps = smp.sample_p_zs_given_zt(torch.from_numpy(s_arr), torch.from_numpy(t_arr), t(z), t(xp), lm, pm, noise=t(noise), n_samples=B)
    gam = O.gamma_table(500, 5.0e-4)
    want_z, want_p = O.sample_p_zs_given_zt(z, xp, eps.cpu().numpy(), noise, np.full(B, gam[249]), np.full(B, gam[250]),
                                            c['lig_mask'], c['pocket_mask'])
    assert np.abs(zs.cpu().numpy() - want_z).max() < 2e-5
    assert np.abs(xps.cpu().numpy() - want_p).max() < 2e-5
    assert xps.shape[1] == 3 + 20 and zs.shape[1] == 3 + 10
    n_p = int((c['pocket_mask'] == 0).sum())
    pocket = {'x': torch.from_numpy(c['xh_pocket'][:, :3].copy()), 'one_hot': torch.from_numpy(c['xh_pocket'][:, 3:].copy()),
              'size': torch.tensor([n_p] * B), 'mask': torch.from_numpy(c['pocket_mask'])}
    sizes = np.bincount(c['lig_mask'])
    xh, xpo, lmo, pmo = smp.sample_given_pocket(pocket, sizes, timesteps=4)
    assert xh.shape == (len(c['lig_mask']), 13) and xpo.shape == (len(c['pocket_mask']), 23)
    assert torch.isfinite(xh).all() and torch.isfinite(xpo).all()


@pytest.mark.parametrize('name', sorted(CASES))
def test_config_from_checkpoint_recovers_configuration(name):
    """A Lightning checkpoint of the reference (``ddpm.dynamics.*`` state + ``hyper_parameters`` with ``egnn_params`` /
    ``diffusion_params`` Namespaces, lightning_modules.py:32-57) gives back the configuration the weights were made for."""
    from argparse import Namespace
    from dataclasses import asdict
    import torch
    from diffndm_b200.generate import config_from_checkpoint, state_dict_from_checkpoint
    cfg = case_config(name)
    W = random_init(cfg, 7, 0.3)
    egnn = Namespace(edge_cutoff_ligand=cfg.edge_cutoff_ligand, edge_cutoff_pocket=cfg.edge_cutoff_pocket,
                     edge_cutoff_interaction=cfg.edge_cutoff_interaction, norm_constant=cfg.norm_constant,
                     normalization_factor=cfg.normalization_factor, attention=True, tanh=True, reflection_equivariant=False,
                     inv_sublayers=1)
    diff = Namespace(diffusion_steps=500, diffusion_noise_schedule='polynomial_2', diffusion_noise_precision=5.0e-4,
                     normalize_factors=[1, 4])
    ckpt = {'state_dict': {'ddpm.dynamics.' + k: torch.from_numpy(v) for k, v in W.items()} | {'ddpm.gamma.gamma': torch.zeros(501)},
            'hyper_parameters': {'egnn_params': egnn, 'diffusion_params': diff, 'mode': 'pocket_conditioning',
                                 'node_histogram': np.ones((30, 400))}}
    state, hp = state_dict_from_checkpoint(ckpt)
    assert set(state) == set(W)
    got, rep, kw = config_from_checkpoint(state, hp)
    assert asdict(got) == asdict(cfg)
    assert rep == ('CA' if cfg.residue_nf == 20 else 'full-atom')
    assert kw == dict(timesteps=500, noise_schedule='polynomial_2', noise_precision=5.0e-4, norm_values=(1, 4))
    with pytest.raises(NotImplementedError):
        config_from_checkpoint(state, {'mode': 'joint'})


def _write_residue_pdb(path, centers, resnames, lig_xyz):
    """One residue per centre: N / CA / C / O around it (CA exactly on the centre), plus a hetero ligand residue 900."""
    offs = np.array([[-1.2, 0.4, 0.0], [0.0, 0.0, 0.0], [1.3, 0.5, 0.0], [1.9, 1.5, 0.4]], np.float32)
    lines, k = [], 0
    for r, (c, rn) in enumerate(zip(centers, resnames)):
        for name, el, o in zip(['N', 'CA', 'C', 'O'], ['N', 'C', 'C', 'O'], offs):
            k += 1
            p = c + o
            lines.append(f"ATOM  {k:5d}  {name:<3s} {rn} A{r + 1:4d}    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}  1.00  0.00          {el:>2s}")
    for j, p in enumerate(lig_xyz):
        k += 1
        lines.append(f"HETATM{k:5d}  C{j:<2d} LIG A 900    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}  1.00  0.00           C")
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\nEND\n')


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(CASES))
def test_checkpoint_file_to_sdf(name, tmp_path):
    """The script path for the other configurations: a Lightning-layout checkpoint FILE of the reference -> ``from_checkpoint``
    (sizes off the weight shapes, cutoffs / schedule / size histogram off ``hyper_parameters``) -> PDB file -> SDF file on the
    engine.  The loaded denoiser must be the one the weights describe: its forward equals the engine built directly from
    (config, weights), bit for bit."""
    from argparse import Namespace
    import torch
    from diffndm_b200 import output
    from diffndm_b200.engine import B200EGNNDynamics
    from diffndm_b200.generate import LigandGenerator
    cfg = case_config(name)
    W = random_init(cfg, 7, 0.3)
    hist = np.zeros((30, 400))
    hist[8:16, :] = 1.0
    ckpt = {'state_dict': {'ddpm.dynamics.' + k: torch.from_numpy(v) for k, v in W.items()},
            'hyper_parameters': {'egnn_params': Namespace(edge_cutoff_ligand=cfg.edge_cutoff_ligand, edge_cutoff_pocket=cfg.edge_cutoff_pocket,
                                                          edge_cutoff_interaction=cfg.edge_cutoff_interaction, norm_constant=1.0,
                                                          normalization_factor=100.0, attention=True, tanh=True,
                                                          reflection_equivariant=False, inv_sublayers=1),
                                 'diffusion_params': Namespace(diffusion_steps=500, diffusion_noise_schedule='polynomial_2',
                                                               diffusion_noise_precision=5.0e-4, normalize_factors=[1, 4]),
                                 'mode': 'pocket_conditioning', 'node_histogram': hist,
                                 'pocket_representation': 'CA' if cfg.residue_nf == 20 else 'full-atom'}}
    path = tmp_path / 'model.ckpt'
    torch.save(ckpt, path)
    gen = LigandGenerator.from_checkpoint(path)
    assert gen.pocket_representation == ('CA' if cfg.residue_nf == 20 else 'full-atom')
    assert gen.ddpm.dynamics.cfg.hidden_nf == cfg.hidden_nf and gen.ddpm.dynamics.cfg.edge_embedding_dim == cfg.edge_embedding_dim

    rng = np.random.default_rng(11)
    centers = (rng.normal(size=(40, 3)) * 5.0 + np.array([20.0, -8.0, 3.0])).astype(np.float32)
    names = ['ALA', 'GLY', 'SER', 'LEU', 'ASP', 'LYS', 'PHE', 'CYS']
    pdb = tmp_path / 'pocket.pdb'
    _write_residue_pdb(pdb, np.round(centers, 3), [names[i % 8] for i in range(40)],
                       centers.mean(0, keepdims=True) + np.array([[0, 0, 0], [1.4, 0, 0]], np.float32))
    torch.manual_seed(3)
    torch.cuda.manual_seed(3)
    n = gen.generate_to_sdf(str(pdb), tmp_path / 'out.sdf', n_samples=4, batch_size=2, ref_ligand='A:900', timesteps=10)
    mols = output.read_sdf(tmp_path / 'out.sdf')
    assert n == 4 and len(mols) == 4 and all(np.isfinite(m.positions).all() for m in mols)

    # same function as the engine built directly from (config, weights)
    c = {k.split('/', 1)[1]: FX[k] for k in FX.files if k.startswith(name + '/')}
    dev = torch.device('cuda', 0)
    direct = B200EGNNDynamics(cfg, W).eval()
    args = [torch.from_numpy(c[k]).to(dev) for k in ('xh_lig', 'xh_pocket', 't', 'lig_mask', 'pocket_mask')]
    a, _ = gen.ddpm.dynamics(*args)
    b, _ = direct(*args)
    assert torch.equal(a, b)
