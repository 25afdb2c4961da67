"""CPU oracle for the DiffNDM / DiffSBDD denoiser hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference algorithm.  It is the
*checker* for the CUDA engine: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package (``diffndm_b200``) never imports anything under ``oracle/``.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md §8c).  The oracle is therefore pinned against outputs of the
reference itself, produced in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference`` unmodified behind shims) and committed under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.

Every function cites the reference file:line it restates (paths relative to the
reference repository root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np


# ----------------------------------------------------------------------------
# configuration  (configs/crossdock_fullatom_cond.yml:36-58, lightning_modules.py:138-174)
# ----------------------------------------------------------------------------
@dataclass
class OracleConfig:
    atom_nf: int = 10
    residue_nf: int = 10
    n_dims: int = 3
    joint_nf: int = 128
    hidden_nf: int = 256
    n_layers: int = 6
    edge_cutoff_ligand: Optional[float] = None
    edge_cutoff_pocket: Optional[float] = 5.0
    edge_cutoff_interaction: Optional[float] = 5.0
    norm_constant: float = 1.0
    normalization_factor: float = 100.0
    coords_range: float = 15.0          # egnn_new.py:218 passes the undivided value
    timesteps: int = 500
    noise_precision: float = 5.0e-4
    norm_values: tuple = (1.0, 4.0)
    norm_biases: tuple = (None, 0.0)


# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------
def silu(x):
    # torch.nn.SiLU: x * sigmoid(x)
    with np.errstate(over='ignore'):
        return x / (1.0 + np.exp(-x))


def sigmoid(x):
    with np.errstate(over='ignore'):
        return 1.0 / (1.0 + np.exp(-x))


def linear(x, w, b=None):
    # torch.nn.Linear: y = x W^T + b ; w is [out, in]
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def segment_sum(data, seg, n):
    """unsorted_segment_sum body, egnn_new.py:319-326 (scatter_add_ into zeros)."""
    out = np.zeros((n,) + data.shape[1:], dtype=data.dtype)
    np.add.at(out, seg, data)
    return out


def segment_mean(data, seg, n):
    """torch_scatter.scatter_mean / unsorted_segment_sum(aggregation='mean'), egnn_new.py:328-334."""
    s = segment_sum(data, seg, n)
    cnt = np.bincount(seg, minlength=n).astype(data.dtype)
    cnt[cnt == 0] = 1
    return s / cnt[:, None]


# ----------------------------------------------------------------------------
# (a2) edge construction -- dynamics.py:169-187
# ----------------------------------------------------------------------------
def pair_d2(xa, xb):
    """Squared distance, fp32, direct differences, ((dx*dx + dy*dy) + dz*dz), one rounding per op.

    The reference uses torch.cdist(...) <= cutoff (dynamics.py:174-181).  cdist's
    arithmetic is backend dependent (matmul expansion on CPU for >25 rows); the
    oracle and the CUDA kernel both use THIS definition and compare d2 <= cutoff^2,
    and the golden test reports every pair within 1e-4 A of the cutoff.
    """
    xa = xa.astype(np.float32)
    xb = xb.astype(np.float32)
    d = xa[:, None, :] - xb[None, :, :]
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def get_edges(mask_l, mask_p, x_l, x_p, cfg: OracleConfig):
    """dynamics.py:169-187.  Returns int64 edges[2, E] in the reference's order:
    row-major over the block adjacency [[ll, lp], [pl, pp]] == sorted by (row, col),
    self loops included (dynamics.py:183-185 / torch.where)."""
    mask_l = np.asarray(mask_l)
    mask_p = np.asarray(mask_p)
    n_l, n_p = len(mask_l), len(mask_p)
    rows, cols = [], []
    c_l = None if cfg.edge_cutoff_ligand is None else np.float32(cfg.edge_cutoff_ligand) ** 2
    c_p = None if cfg.edge_cutoff_pocket is None else np.float32(cfg.edge_cutoff_pocket) ** 2
    c_i = None if cfg.edge_cutoff_interaction is None else np.float32(cfg.edge_cutoff_interaction) ** 2
    samples = np.union1d(np.unique(mask_l), np.unique(mask_p))
    per_row_cols = [None] * (n_l + n_p)
    for b in samples:
        il = np.nonzero(mask_l == b)[0]
        ip = np.nonzero(mask_p == b)[0]
        xl, xp = x_l[il], x_p[ip]
        a_ll = np.ones((len(il), len(il)), bool)
        if c_l is not None:
            a_ll &= pair_d2(xl, xl) <= c_l
        a_pp = np.ones((len(ip), len(ip)), bool)
        if c_p is not None:
            a_pp &= pair_d2(xp, xp) <= c_p
        a_lp = np.ones((len(il), len(ip)), bool)
        if c_i is not None:
            a_lp &= pair_d2(xl, xp) <= c_i
        for k, i in enumerate(il):
            per_row_cols[i] = np.concatenate([il[a_ll[k]], n_l + ip[a_lp[k]]])
        for k, i in enumerate(ip):
            per_row_cols[n_l + i] = np.concatenate([il[a_lp[:, k]], n_l + ip[a_pp[k]]])
    for i, c in enumerate(per_row_cols):
        if c is None:
            continue
        c = np.sort(c)
        rows.append(np.full(len(c), i, np.int64))
        cols.append(c.astype(np.int64))
    if not rows:
        return np.zeros((2, 0), np.int64)
    return np.stack([np.concatenate(rows), np.concatenate(cols)])


# ----------------------------------------------------------------------------
# (a5) geometry -- egnn_new.py:296-316
# ----------------------------------------------------------------------------
def coord2diff(x, row, col, norm_constant):
    """egnn_new.py:296-302."""
    d = x[row] - x[col]
    radial = np.sum(d * d, axis=1, keepdims=True)
    norm = np.sqrt(radial + 1e-8)
    return radial, d / (norm + norm_constant)


def coord2cross(x, row, col, batch_mask, norm_constant):
    """egnn_new.py:305-316 (mean over ALL nodes of the sample, ligand+pocket)."""
    nb = int(batch_mask.max()) + 1
    mean = segment_mean(x, batch_mask, nb)
    a = x[row] - mean[batch_mask[row]]
    b = x[col] - mean[batch_mask[col]]
    cross = np.cross(a, b)
    norm = np.linalg.norm(cross, axis=1, keepdims=True)
    return cross / (norm + norm_constant)


# ----------------------------------------------------------------------------
# (a6,a7) GCL / EquivariantUpdate -- egnn_new.py:31-66, 96-132
# ----------------------------------------------------------------------------
def gcl_forward(W, p, h, row, col, edge_attr, cfg):
    """GCL.forward, egnn_new.py:59-66 with edge_model :31-47 and node_model :49-57."""
    inp = np.concatenate([h[row], h[col], edge_attr], axis=1)
    m = silu(linear(inp, W[p + 'edge_mlp.0.weight'], W[p + 'edge_mlp.0.bias']))
    mij = silu(linear(m, W[p + 'edge_mlp.2.weight'], W[p + 'edge_mlp.2.bias']))
    att = sigmoid(linear(mij, W[p + 'att_mlp.0.weight'], W[p + 'att_mlp.0.bias']))
    edge_feat = mij * att
    agg = segment_sum(edge_feat, row, h.shape[0]) / h.dtype.type(cfg.normalization_factor)
    cat = np.concatenate([h, agg], axis=1)
    out = linear(silu(linear(cat, W[p + 'node_mlp.0.weight'], W[p + 'node_mlp.0.bias'])),
                 W[p + 'node_mlp.2.weight'], W[p + 'node_mlp.2.bias'])
    return h + out, agg


def equiv_forward(W, p, h, x, row, col, coord_diff, coord_cross, edge_attr, update_mask, cfg):
    """EquivariantUpdate.coord_model, egnn_new.py:96-123 (tanh=True, reflection_equiv=False)."""
    inp = np.concatenate([h[row], h[col], edge_attr], axis=1)

    def head(q):
        a = silu(linear(inp, W[p + q + '.0.weight'], W[p + q + '.0.bias']))
        a = silu(linear(a, W[p + q + '.2.weight'], W[p + q + '.2.bias']))
        return linear(a, W[p + q + '.4.weight'])

    rng = x.dtype.type(cfg.coords_range)
    phi = np.tanh(head('coord_mlp')) * rng
    psi = np.tanh(head('cross_product_mlp')) * rng
    trans = coord_diff * phi + coord_cross * psi
    agg = segment_sum(trans, row, x.shape[0]) / x.dtype.type(cfg.normalization_factor)
    if update_mask is not None:
        agg = update_mask * agg
    return x + agg, phi, psi


# ----------------------------------------------------------------------------
# (a1,a3,a4) EGNNDynamics.forward -- dynamics.py:87-167, egnn_new.py:225-244
# ----------------------------------------------------------------------------
def dynamics_forward(W: Dict[str, np.ndarray], xh_atoms, xh_residues, t, mask_atoms, mask_residues,
                     cfg: OracleConfig, dtype=np.float32, edges=None, trace: Optional[dict] = None):
    """EGNNDynamics.forward (mode='egnn_dynamics', condition_time, update_pocket_coords=False; with the edge-type
    embedding when the table carries ``edge_embedding.weight``) -- dynamics.py:87-167.

    ``W`` uses the reference state_dict key names.  ``dtype=np.float64``
    gives the "truth" run used to set tolerances.  ``edges`` may be injected (int64 [2,E]).
    Returns (out_lig [N_l, 3+atom_nf], out_pocket [N_p, 3+residue_nf]).
    """
    W = {k: v.astype(dtype) for k, v in W.items()}
    xh_atoms = np.asarray(xh_atoms, dtype)
    xh_residues = np.asarray(xh_residues, dtype)
    mask_atoms = np.asarray(mask_atoms, np.int64)
    mask_residues = np.asarray(mask_residues, np.int64)
    nd = cfg.n_dims
    n_l = len(mask_atoms)

    x_atoms, h_atoms = xh_atoms[:, :nd], xh_atoms[:, nd:]
    x_res, h_res = xh_residues[:, :nd], xh_residues[:, nd:]

    # dynamics.py:96-97
    h_atoms = linear(silu(linear(h_atoms, W['atom_encoder.0.weight'], W['atom_encoder.0.bias'])),
                     W['atom_encoder.2.weight'], W['atom_encoder.2.bias'])
    h_res = linear(silu(linear(h_res, W['residue_encoder.0.weight'], W['residue_encoder.0.bias'])),
                   W['residue_encoder.2.weight'], W['residue_encoder.2.bias'])

    # dynamics.py:100-111
    x = np.concatenate([x_atoms, x_res], axis=0)
    h = np.concatenate([h_atoms, h_res], axis=0)
    mask = np.concatenate([mask_atoms, mask_residues])
    t = np.asarray(t, dtype)
    if t.size == 1:
        h_time = np.full((h.shape[0], 1), t.reshape(-1)[0], dtype)
    else:
        h_time = t.reshape(-1, 1)[mask]
    h = np.concatenate([h, h_time], axis=1)

    # dynamics.py:114  (edge decisions are ALWAYS taken in fp32, see pair_d2)
    if edges is None:
        edges = get_edges(mask_atoms, mask_residues, x_atoms.astype(np.float32),
                          x_res.astype(np.float32), cfg)
    row, col = edges[0], edges[1]
    assert np.all(mask[row] == mask[col])          # dynamics.py:115

    # dynamics.py:130-132  update_coords_mask = [1..1 | 0..0]
    update_mask = np.concatenate([np.ones(n_l, dtype), np.zeros(len(mask_residues), dtype)])[:, None]

    # EGNN.forward, egnn_new.py:225-244
    r0, _ = coord2diff(x, row, col, 1)                                      # :228 (default norm_constant)
    if 'edge_embedding.weight' in W:                                        # dynamics.py:118-127 (edge_embedding_dim, moad configs)
        # 0: ligand-pocket, 1: ligand-ligand, 2: pocket-pocket; the learned embedding rides behind r0 (egnn_new.py:230-231)
        etype = np.where((row < n_l) & (col < n_l), 1, np.where((row >= n_l) & (col >= n_l), 2, 0))
        r0 = np.concatenate([r0, W['edge_embedding.weight'][etype]], axis=1)
    h = linear(h, W['egnn.embedding.weight'], W['egnn.embedding.bias'])     # :233
    x_cur = x
    if trace is not None:
        trace['edges'] = edges
        trace['h_embed'] = h.copy()
    for i in range(cfg.n_layers):
        p = f'egnn.e_block_{i}.'
        # EquivariantBlock.forward, egnn_new.py:163-184
        radial, coord_diff = coord2diff(x_cur, row, col, cfg.norm_constant)
        coord_cross = coord2cross(x_cur, row, col, mask, cfg.norm_constant)
        edge_attr = np.concatenate([radial, r0], axis=1)                    # :174
        h, _ = gcl_forward(W, p + 'gcl_0.', h, row, col, edge_attr, cfg)    # :176
        x_cur, phi, psi = equiv_forward(W, p + 'gcl_equiv.', h, x_cur, row, col, coord_diff,
                                        coord_cross, edge_attr, update_mask, cfg)   # :178
        if trace is not None:
            trace[f'h_{i}'] = h.copy()
            trace[f'x_{i}'] = x_cur.copy()
            trace[f'phi_{i}'] = phi[:, 0].copy()
            trace[f'psi_{i}'] = psi[:, 0].copy()
    h = linear(h, W['egnn.embedding_out.weight'], W['egnn.embedding_out.bias'])     # :241

    vel = x_cur - x                                                          # dynamics.py:136
    h_final = h[:, :-1]                                                      # :147-149
    h_fa = linear(silu(linear(h_final[:n_l], W['atom_decoder.0.weight'], W['atom_decoder.0.bias'])),
                  W['atom_decoder.2.weight'], W['atom_decoder.2.bias'])
    h_fr = linear(silu(linear(h_final[n_l:], W['residue_decoder.0.weight'], W['residue_decoder.0.bias'])),
                  W['residue_decoder.2.weight'], W['residue_decoder.2.bias'])
    if np.any(np.isnan(vel)):                                                # :155-159 (eval mode)
        raise ValueError("NaN detected in EGNN output")
    return (np.concatenate([vel[:n_l], h_fa], axis=1),
            np.concatenate([vel[n_l:], h_fr], axis=1))


# ----------------------------------------------------------------------------
# (a11) noise schedule -- en_diffusion.py:1119-1195
# ----------------------------------------------------------------------------
def clip_noise_schedule(alphas2, clip_value=0.001):
    """en_diffusion.py:1119-1133."""
    alphas2 = np.concatenate([np.ones(1), alphas2], axis=0)
    alphas_step = alphas2[1:] / alphas2[:-1]
    alphas_step = np.clip(alphas_step, a_min=clip_value, a_max=1.)
    return np.cumprod(alphas_step, axis=0)


def polynomial_schedule(timesteps, s=1e-4, power=3.):
    """en_diffusion.py:1146-1160."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas2 = (1 - np.power(x / steps, power)) ** 2
    alphas2 = clip_noise_schedule(alphas2, clip_value=0.001)
    precision = 1 - 2 * s
    return precision * alphas2 + s


def gamma_table(timesteps=500, precision=5.0e-4, power=2.0):
    """PredefinedNoiseSchedule.__init__ for 'polynomial_<power>', en_diffusion.py:1163-1191.
    Returns float32 gamma[T+1]; lookup is gamma[round(t*T)] (:1193-1195)."""
    alphas2 = polynomial_schedule(timesteps, s=precision, power=power)
    sigmas2 = 1 - alphas2
    return (-(np.log(alphas2) - np.log(sigmas2))).astype(np.float32)


def _softplus(x):
    return np.log1p(np.exp(-np.abs(x))) + np.maximum(x, 0)


def _logsigmoid(x):
    return -_softplus(-x)


def step_scalars(gamma_s, gamma_t):
    """sigma_and_alpha_t_given_s (en_diffusion.py:83-108) + sigma() (:870-873), fp32.
    Returns dict of per-sample scalars."""
    gs = np.asarray(gamma_s, np.float32)
    gt = np.asarray(gamma_t, np.float32)
    sigma2_ts = -np.expm1(_softplus(gs) - _softplus(gt))
    alpha_ts = np.exp(np.float32(0.5) * (_logsigmoid(-gt) - _logsigmoid(-gs)))
    sigma_ts = np.sqrt(sigma2_ts)
    sigma_s = np.sqrt(sigmoid(gs))
    sigma_t = np.sqrt(sigmoid(gt))
    return dict(sigma2_ts=sigma2_ts.astype(np.float32), alpha_ts=alpha_ts.astype(np.float32),
                sigma_ts=sigma_ts.astype(np.float32), sigma_s=sigma_s.astype(np.float32),
                sigma_t=sigma_t.astype(np.float32))


# ----------------------------------------------------------------------------
# (a9,a10) sampler step -- conditional_model.py:483-540, 165-186, 1793-1801
# ----------------------------------------------------------------------------
def remove_mean_batch(x_lig, x_pocket, lig_mask, pocket_mask, nb):
    """conditional_model.py:1793-1801: subtract the per-sample LIGAND mean from both."""
    mean = segment_mean(x_lig, lig_mask, nb)
    return x_lig - mean[lig_mask], x_pocket - mean[pocket_mask]


def sample_p_zs_given_zt(z_lig, xh_pocket, eps_lig, noise, gamma_s, gamma_t, lig_mask, pocket_mask):
    """conditional_model.py:483-540 with the dynamics output ``eps_lig`` and the Gaussian draw
    ``noise`` (sample_gaussian, :172-174) injected.  gamma_s/gamma_t: per-sample [B]."""
    z_lig = np.asarray(z_lig, np.float32)
    xh_pocket = np.asarray(xh_pocket, np.float32)
    nb = len(gamma_s)
    sc = step_scalars(gamma_s, gamma_t)
    a_ts = sc['alpha_ts'][lig_mask][:, None]
    c_eps = (sc['sigma2_ts'] / sc['alpha_ts'] / sc['sigma_t'])[lig_mask][:, None]
    mu = z_lig / a_ts - c_eps * np.asarray(eps_lig, np.float32)              # :524-526
    sigma = (sc['sigma_ts'] * sc['sigma_s'] / sc['sigma_t'])[lig_mask][:, None]   # :529
    out = mu + sigma * np.asarray(noise, np.float32)                         # :176
    xh_p = xh_pocket.copy()
    out[:, :3], xh_p[:, :3] = remove_mean_batch(out[:, :3], xh_pocket[:, :3], lig_mask, pocket_mask, nb)
    return out.astype(np.float32), xh_p.astype(np.float32)


def sample_p_xh_given_z0(z0_lig, xh0_pocket, eps0_lig, noise, gamma_0, lig_mask, pocket_mask, cfg):
    """conditional_model.py:136-160 with the t=0 dynamics output and noise injected.
    Returns x_lig [N_l,3], atom type index [N_l] (argmax; one-hot in the reference), x_pocket, h_pocket."""
    g0 = np.asarray(gamma_0, np.float32)
    nb = len(g0)
    sigma_x = np.exp(np.float32(0.5) * g0)                  # SNR(-0.5*gamma_0), en_diffusion.py:880-883
    sigma0 = np.sqrt(sigmoid(g0))
    alpha0 = np.sqrt(sigmoid(-g0))
    z0 = np.asarray(z0_lig, np.float32)
    mu = (1.0 / alpha0)[lig_mask][:, None] * (z0 - sigma0[lig_mask][:, None] * np.asarray(eps0_lig, np.float32))
    out = mu + sigma_x[lig_mask][:, None] * np.asarray(noise, np.float32)
    xp = np.asarray(xh0_pocket, np.float32).copy()
    out[:, :3], xp[:, :3] = remove_mean_batch(out[:, :3], xp[:, :3], lig_mask, pocket_mask, nb)
    x_lig = out[:, :3] * np.float32(cfg.norm_values[0])
    h_lig = z0[:, 3:] * np.float32(cfg.norm_values[1]) + np.float32(cfg.norm_biases[1])
    x_pocket = xp[:, :3] * np.float32(cfg.norm_values[0])
    h_pocket = xp[:, 3:] * np.float32(cfg.norm_values[1]) + np.float32(cfg.norm_biases[1])
    return x_lig, np.argmax(h_lig, axis=1), x_pocket, h_pocket


def x0_lookahead_z0(z_t, eps_t, gamma_t, lig_mask):
    """my_to_x0 first half, conditional_model.py:457-464: z0 = (z_t - sigma_t eps)/alpha_t."""
    gt = np.asarray(gamma_t, np.float32)
    alpha_t = np.exp(np.float32(0.5) * _logsigmoid(-gt))
    sigma_t = np.sqrt(sigmoid(gt))
    return (np.asarray(z_t, np.float32) - sigma_t[lig_mask][:, None] * np.asarray(eps_t, np.float32)) \
        / alpha_t[lig_mask][:, None]


# ----------------------------------------------------------------------------
# (a13) SPSA update term -- conditional_model.py:738-813
# ----------------------------------------------------------------------------
def spsa_update(z_lig, xh_pocket, perturbations, f_plus, f_minus, lig_mask, pocket_mask,
                guidance_scale=1e-3, zeta_div=1e-4):
    """my_gradient_for_molecule (:738-759) averaged over k draws + the update at :801-812.
    perturbations: [k, N_l, 3]; f_plus/f_minus: [k, B] host rewards.  NB the divisor is the
    hard-coded 2*1e-4 of :799, not the perturbation scale."""
    z = np.asarray(z_lig, np.float32).copy()
    k = len(perturbations)
    nb = np.asarray(f_plus).shape[1]
    g = np.zeros_like(z[:, :3])
    for i in range(k):
        dd = ((np.asarray(f_plus[i], np.float32) - np.asarray(f_minus[i], np.float32))
              / np.float32(2 * zeta_div))[lig_mask][:, None]
        g = g + dd * np.asarray(perturbations[i], np.float32)
    g = g / np.float32(k)
    z[:, :3] = z[:, :3] + np.float32(guidance_scale) * g
    xp = np.asarray(xh_pocket, np.float32).copy()
    z[:, :3], xp[:, :3] = remove_mean_batch(z[:, :3], xp[:, :3], lig_mask, pocket_mask, nb)
    return z, xp


# ----------------------------------------------------------------------------
# (a14) ATP selection -- conditional_model.py:1203-1232
# ----------------------------------------------------------------------------
def atp_select(r0, r, s, big_z, big_pocket, big_lig_mask, big_pocket_mask, top_k):
    """mixed = r0*(s/250) + r*(250 - s/250)  [sic, :1203]; global top-k over all candidates (:1205); winners re-batched
    in rank order (:1212-1232).  Ties are broken by the lower candidate index (torch.topk on CPU keeps that order for
    the fixtures used here).  Returns (z_lig, xh_pocket, lig_mask) of the selected candidates."""
    r0 = np.asarray(r0, np.float32)
    r = np.asarray(r, np.float32)
    mixed = r0 * np.float32(s / 250) + r * np.float32(250 - s / 250)
    order = np.argsort(-mixed, kind='stable')[:top_k]
    zs, ps, ms = [], [], []
    for rank, idx in enumerate(order):
        nm = big_lig_mask == idx
        zs.append(big_z[nm])
        ps.append(big_pocket[big_pocket_mask == idx])
        ms.append(np.full(int(nm.sum()), rank, np.int64))
    return np.concatenate(zs), np.concatenate(ps), np.concatenate(ms), order


# ----------------------------------------------------------------------------
# inpainting (RePaint resampling) -- conditional_model.py:1491-1790, 188-215, 470-481
# ----------------------------------------------------------------------------
def noised_representation(xh_lig, xh_pocket, noise, gamma, lig_mask, pocket_mask):
    """q(z_t | x): alpha_t x + sigma_t eps, then the ligand-COM projection (conditional_model.py:188-215)."""
    g = np.asarray(gamma, np.float32)
    nb = len(g)
    alpha = np.sqrt(sigmoid(-g))[lig_mask][:, None]
    sigma = np.sqrt(sigmoid(g))[lig_mask][:, None]
    z = (alpha * np.asarray(xh_lig, np.float32) + sigma * np.asarray(noise, np.float32)).astype(np.float32)
    xp = np.asarray(xh_pocket, np.float32).copy()
    z[:, :3], xp[:, :3] = remove_mean_batch(z[:, :3], xp[:, :3], lig_mask, pocket_mask, nb)
    return z, xp


def sample_p_zt_given_zs(zs_lig, xh_pocket, noise, gamma_t, gamma_s, lig_mask, pocket_mask):
    """Forward (re-noising) move of the resampling loop: z_t = alpha_{t|s} z_s + sigma_{t|s} eps, COM-projected
    (conditional_model.py:470-481 via sample_normal_zero_com :165-186)."""
    sc = step_scalars(gamma_s, gamma_t)
    nb = len(sc['alpha_ts'])
    z = (sc['alpha_ts'][lig_mask][:, None] * np.asarray(zs_lig, np.float32)
         + sc['sigma_ts'][lig_mask][:, None] * np.asarray(noise, np.float32)).astype(np.float32)
    xp = np.asarray(xh_pocket, np.float32).copy()
    z[:, :3], xp[:, :3] = remove_mean_batch(z[:, :3], xp[:, :3], lig_mask, pocket_mask, nb)
    return z, xp


def inpaint_combine(z_known, z_unknown, xh_pocket, lig_fixed, lig_mask, pocket_mask, nb):
    """Move the noised known part onto the COM of the denoised one (both over the FIXED atoms), shift the pocket with it
    and blend (conditional_model.py:1596-1609)."""
    fx = np.asarray(lig_fixed).reshape(-1) > 0
    com_noised = segment_mean(z_known[fx][:, :3], lig_mask[fx], nb)
    com_denoised = segment_mean(z_unknown[fx][:, :3], lig_mask[fx], nb)
    dx = (com_denoised - com_noised).astype(np.float32)
    zk = z_known.copy()
    zk[:, :3] = zk[:, :3] + dx[lig_mask]
    xp = xh_pocket.copy()
    xp[:, :3] = xp[:, :3] + dx[pocket_mask]
    f = fx.astype(np.float32)[:, None]
    return (zk * f + z_unknown * (1 - f)).astype(np.float32), xp


def inpaint(W, lig_x, lig_onehot, lig_mask, pocket_x, pocket_onehot, pocket_mask, lig_fixed, noise, timesteps,
            resamplings, cfg: OracleConfig = OracleConfig(), T=500, dtype=np.float32, eps_fn=None):
    """ConditionalDDPM.inpaint (conditional_model.py:1491-1790) with center='ligand', svdd=0 and every Gaussian draw
    injected in the reference's order (``noise`` [n_draws, N_l, 3+atom_nf]); the SPSA window (12 <= s <= 16) is outside
    the fixtures' range.  Returns (xh_lig with one-hot features, xh_pocket) in Angstrom like the reference."""
    gam = gamma_table(T)
    nb = int(lig_mask.max()) + 1
    nv0, nv1, nb1 = np.float32(cfg.norm_values[0]), np.float32(cfg.norm_values[1]), np.float32(cfg.norm_biases[1])
    lx = np.asarray(lig_x, np.float32) / nv0
    lh = (np.asarray(lig_onehot, np.float32) - nb1) / nv1
    xh0_pocket = np.concatenate([np.asarray(pocket_x, np.float32) / nv0,
                                 (np.asarray(pocket_onehot, np.float32) - nb1) / nv1], 1)
    com_pocket_0 = segment_mean(xh0_pocket[:, :3], pocket_mask, nb)
    xh_ligand = np.concatenate([lx, lh], 1)
    fx = np.asarray(lig_fixed).reshape(-1) > 0
    mean_known = segment_mean(lx[fx], lig_mask[fx], nb)
    it = iter(noise)
    mu = np.concatenate([mean_known, np.zeros((nb, lh.shape[1]), np.float32)], 1)[lig_mask]
    z = (mu + next(it)).astype(np.float32)
    xp = xh0_pocket.copy()
    z[:, :3], xp[:, :3] = remove_mean_batch(z[:, :3], xh0_pocket[:, :3], lig_mask, pocket_mask, nb)
    look = lambda tt: gam[int(round(tt * T))]
    if eps_fn is None:     # ``eps_fn(z, xp, t[B,1], lig_mask, pocket_mask) -> eps_lig`` replaces the plain denoiser call (tests wrap it)
        eps_fn = lambda z_, xp_, t_, lm_, pm_: dynamics_forward(W, z_, xp_, t_, lm_, pm_, cfg, dtype=dtype)[0]
    for s in reversed(range(timesteps)):
        for u in range(resamplings):
            ts, tt = s / timesteps, (s + 1) / timesteps
            g_s = np.full(nb, look(ts), np.float32)
            g_t = np.full(nb, look(tt), np.float32)
            eps = eps_fn(z, xp, np.full((nb, 1), tt, np.float32), lig_mask, pocket_mask)
            z_unknown, xp = sample_p_zs_given_zt(z, xp, eps.astype(np.float32), next(it), g_s, g_t, lig_mask, pocket_mask)
            com_pocket = segment_mean(xp[:, :3], pocket_mask, nb)
            xh_ligand[:, :3] = lx + (com_pocket - com_pocket_0)[lig_mask]
            z_known, xp = noised_representation(xh_ligand, xp, next(it), g_s, lig_mask, pocket_mask)
            z, xp = inpaint_combine(z_known, z_unknown, xp, lig_fixed, lig_mask, pocket_mask, nb)
            if u < resamplings - 1:
                z, xp = sample_p_zt_given_zs(z, xp, next(it), g_t, g_s, lig_mask, pocket_mask)
    g0 = np.full(nb, gam[0], np.float32)
    eps0 = eps_fn(z, xp, np.zeros((nb, 1), np.float32), lig_mask, pocket_mask)
    x_l, t_l, x_p, h_p = sample_p_xh_given_z0(z, xp, eps0.astype(np.float32), next(it), g0, lig_mask, pocket_mask, cfg)
    onehot = np.eye(lh.shape[1], dtype=np.float32)[t_l]
    return np.concatenate([x_l, onehot], 1), np.concatenate([x_p, h_p], 1), z, xp


# ----------------------------------------------------------------------------
# bond perception: analysis/molecule_builder.py:30-55 (get_bond_order_batch), :100-116 (make_mol_edm)
# ----------------------------------------------------------------------------
def bond_orders(x, atom_types, mol_mask, bonds1, bonds2, bonds3, margins=(3.0, 2.0, 1.0)):
    """Per molecule, the directed lower-triangular bond-order matrix E of make_mol_edm (molecule_builder.py:107-113):
    E[i, j] (i > j) = 3 / 2 / 1 / 0 by the distance tables (pm) + margins, later rules overwriting earlier ones exactly
    like get_bond_order_batch (:43-53).  Distances are fp32: 100 * sqrt((dx*dx + dy*dy) + dz*dz), one rounding per
    operation (torch.cdist's arithmetic is backend dependent; the fixtures count the pairs near a threshold).
    Returns a list of int8 [n_b, n_b] matrices, the per-atom valence (sum of the symmetrised orders) and per molecule
    (n_bonds, n_components, largest_component)."""
    x = np.asarray(x, np.float32)
    t = np.asarray(atom_types, np.int64)
    b1, b2, b3 = (np.asarray(b, np.float32) for b in (bonds1, bonds2, bonds3))
    m1, m2, m3 = (np.float32(m) for m in margins)
    nb = int(mol_mask.max()) + 1
    mats, stats = [], []
    valence = np.zeros(len(t), np.int32)
    for b in range(nb):
        idx = np.nonzero(mol_mask == b)[0]
        xi, ti = x[idx], t[idx]
        d = xi[:, None, :] - xi[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        dist = np.float32(100.0) * np.sqrt(d2.astype(np.float32))
        ta, tb = ti[:, None], ti[None, :]
        e = np.zeros(dist.shape, np.int8)
        e[dist < b1[ta, tb] + m1] = 1
        e[dist < b2[ta, tb] + m2] = 2
        e[dist < b3[ta, tb] + m3] = 3
        e = np.tril(e, -1)
        mats.append(e)
        sym = e.astype(np.int32) + e.astype(np.int32).T
        valence[idx] = sym.sum(1)
        # connected components of the bond graph (process_molecule's largest_frag works on these fragments)
        n = len(idx)
        label = np.arange(n)
        adj = sym > 0
        changed = True
        while changed:
            changed = False
            for i in range(n):
                m = min(label[i], label[adj[i]].min()) if adj[i].any() else label[i]
                if m < label[i]:
                    label[i] = m
                    changed = True
        _, counts = np.unique(label, return_counts=True)
        stats.append((int((e > 0).sum()), len(counts), int(counts.max())))
    return mats, valence, np.asarray(stats, np.int32)


# ----------------------------------------------------------------------------
# algorithmic work model
# ----------------------------------------------------------------------------
def reference_flops(n_nodes, n_edges, cfg: OracleConfig = OracleConfig()):
    """Dense reference-formulation FLOPs of one EGNNDynamics.forward."""
    H, de = cfg.hidden_nf, 2
    per_edge_block = 3 * (2 * (2 * H + de) * H + 2 * H * H) + 3 * 2 * H
    per_node = cfg.n_layers * (2 * 2 * H * H + 2 * H * H) + 2 * 2 * (cfg.joint_nf + 1) * H + 2 * 5520
    return n_edges * per_edge_block * cfg.n_layers + n_nodes * per_node
