"""Shared pieces of the guidance (SPSA / ATP) parity fixtures: used by the generator (reference side, build container)
and by the tests (oracle on CPU, engine on the GPU).  Test infrastructure, not product code.

* ``SyntheticScoreDynamics`` wraps a denoiser (the reference's ``EGNNDynamics`` or ``B200EGNNDynamics``) and adds the exact
  score of a point-mass data distribution, eps = eps_net + (z_t - alpha_t x_0) / sigma_t.  Random-init weights cannot
  denoise: without it every reverse trajectory inflates by 1/alpha_T ~ 45x (|z| ~ 1e2) and absolute tolerances in
  Angstrom mean nothing.  With it the states stay O(1..10) A like a trained model's, on both sides identically (fp32
  elementwise torch ops).  x_0 is given relative to the pocket's first atom, so it follows the pocket translation.
* ``geometric_reward``: deterministic stand-in for ``handle_to_mol`` + ``my_reward_for_SPSA / _SVDD`` (RDKit is absent).
* ``NoiseStream``: every Gaussian draw comes from numpy's PCG64 (stable across platforms), so fixtures store the seed and
  the shapes, not the noise.
"""
from __future__ import annotations

import numpy as np
import torch


def polynomial2_gamma(T=500, precision=5.0e-4):
    """gamma table of PredefinedNoiseSchedule('polynomial_2', T, precision) -- en_diffusion.py:1146-1191."""
    steps = T + 1
    x = np.linspace(0, steps, steps)
    alphas2 = (1 - np.power(x / steps, 2.0)) ** 2
    alphas2 = np.concatenate([np.ones(1), alphas2], axis=0)
    alphas_step = np.clip(alphas2[1:] / alphas2[:-1], a_min=0.001, a_max=1.)
    alphas2 = np.cumprod(alphas_step, axis=0)
    alphas2 = (1 - 2 * precision) * alphas2 + precision
    sigmas2 = 1 - alphas2
    return torch.from_numpy(-(np.log(alphas2) - np.log(sigmas2))).float()


class SyntheticScoreDynamics(torch.nn.Module):
    """forward(xh_atoms, xh_residues, t, mask_atoms, mask_residues[, n_samples]) -> (eps_atoms, eps_residues) like
    ``EGNNDynamics.forward`` (dynamics.py:87-167), with the synthetic score added to the ligand part.

    ``x0_rel`` [n, 3 + atom_nf]: data point in normalised units, coordinates relative to the pocket's first atom; row j of
    the batch uses x0_rel[j % n] (a batch is the base batch repeated, or equally sized ligands sharing one pose)."""

    def __init__(self, inner, x0_rel, T=500):
        super().__init__()
        self.inner = inner
        self.register_buffer('x0_rel', torch.as_tensor(x0_rel, dtype=torch.float32), persistent=False)
        self.register_buffer('gamma_tab', polynomial2_gamma(T), persistent=False)
        self.T = T
        self.update_pocket_coords = inner.update_pocket_coords          # read by ConditionalDDPM.__init__
        self.compute_pocket_output = True
        for name in ('engine', 'cfg', 'n_dims'):
            if hasattr(inner, name):
                object.__setattr__(self, name, getattr(inner, name))

    @property
    def check_nan(self):
        return self.inner.check_nan

    @check_nan.setter
    def check_nan(self, v):
        self.inner.check_nan = v

    def forward(self, xh_atoms, xh_residues, t, mask_atoms, mask_residues, **kw):
        if hasattr(self.inner, 'compute_pocket_output'):
            self.inner.compute_pocket_output = self.compute_pocket_output
        eps_l, eps_p = self.inner(xh_atoms, xh_residues, t, mask_atoms, mask_residues, **kw)
        dev = xh_atoms.device
        x0 = self.x0_rel.to(dev)
        gam = self.gamma_tab.to(dev)
        B = int(kw['n_samples']) if kw.get('n_samples') else (int(t.numel()) if t.numel() > 1 else int(mask_atoms.max()) + 1)
        tt = t.to(dev).float().reshape(-1)
        if tt.numel() == 1:
            tt = tt.expand(B)
        g = gam[torch.round(tt * self.T).long()]
        alpha = torch.sqrt(torch.sigmoid(-g))[mask_atoms][:, None]
        sigma = torch.sqrt(torch.sigmoid(g))[mask_atoms][:, None]
        first = torch.searchsorted(mask_residues.contiguous(), torch.arange(B, device=dev))
        target = x0[torch.arange(xh_atoms.shape[0], device=dev) % x0.shape[0]].clone()
        target[:, :3] += xh_residues[first][mask_atoms, :3].to(target.dtype)
        target = target.to(eps_l.dtype)
        return eps_l + (xh_atoms.to(eps_l.dtype) - alpha.to(eps_l.dtype) * target) / sigma.to(eps_l.dtype), eps_p


def score_numpy(eps_l, xh_atoms, xh_residues, t, mask_atoms, mask_residues, x0_rel, gamma_tab, T=500):
    """The same correction in numpy fp32 for the oracle side."""
    B = int(mask_atoms.max()) + 1 if np.ndim(t) == 0 or np.size(t) == 1 else int(np.size(t))
    tt = np.broadcast_to(np.asarray(t, np.float32).reshape(-1), (B,)) if np.size(t) == 1 else np.asarray(t, np.float32).reshape(-1)
    g = np.asarray(gamma_tab, np.float32)[np.round(tt * T).astype(np.int64)]
    sig = lambda v: (1.0 / (1.0 + np.exp(-v.astype(np.float64)))).astype(np.float32)
    alpha = np.sqrt(sig(-g))[mask_atoms][:, None]
    sigma = np.sqrt(sig(g))[mask_atoms][:, None]
    first = np.searchsorted(mask_residues, np.arange(B))
    target = np.asarray(x0_rel, np.float32)[np.arange(len(xh_atoms)) % len(x0_rel)].copy()
    target[:, :3] += xh_residues[first][mask_atoms, :3]
    return (eps_l + (xh_atoms - alpha * target) / sigma).astype(np.float32)


def geometric_reward(x, types, mask):
    """Per-molecule score from coordinates [N,3], atom types [N] and the molecule mask [N] (sorted): compactness, a
    bond-length term and a type term.  Smooth in x, O(1), translation invariant.  Returns a list of python floats."""
    x = np.asarray(x, np.float64)
    types = np.asarray(types, np.int64)
    mask = np.asarray(mask, np.int64)
    out = []
    for b in range(int(mask.max()) + 1 if len(mask) else 0):
        idx = np.nonzero(mask == b)[0]
        xi, ti = x[idx], types[idx]
        c = xi.mean(0)
        rg2 = ((xi - c) ** 2).sum(1).mean()
        d = np.sqrt(((xi[:, None] - xi[None]) ** 2).sum(-1) + 1e-12)
        iu = np.triu_indices(len(idx), 1)
        bond = np.exp(-(d[iu] - 1.5) ** 2).mean() if len(iu[0]) else 0.0
        out.append(float(2.0 * np.exp(-rg2 / 8.0) + 3.0 * bond + 0.2 * np.sin(1.3 * ti + 0.5).mean()))
    return out


def stress_weights(W, factor=20.0):
    """The two radial input columns (r^2 of the current / of the input coordinates) of every edge MLP scaled by ``factor``:
    trained-scale radial weights for the radial stress fixture (hard part 1)."""
    W = {k: v.copy() for k, v in W.items()}
    for k in W:
        if k.endswith('edge_mlp.0.weight') or k.endswith('coord_mlp.0.weight') or k.endswith('cross_product_mlp.0.weight'):
            W[k][:, -2:] *= np.float32(factor)
    return W


class NoiseStream:
    """Replacement for ``torch.randn`` backed by numpy PCG64; records the shape of every draw."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(int(seed))
        self.shapes = []

    def randn(self, *size, device=None, dtype=None, **kw):
        if len(size) == 1 and not isinstance(size[0], int):
            size = tuple(size[0])
        self.shapes.append(tuple(int(s) for s in size))
        a = self.rng.standard_normal(size).astype(np.float32)
        t = torch.from_numpy(a)
        if dtype is not None:
            t = t.to(dtype)
        return t if device is None else t.to(device)

    def draw(self, *size):
        """numpy draw (test side)."""
        self.shapes.append(tuple(int(s) for s in size))
        return self.rng.standard_normal(size).astype(np.float32)


def reference_order_noise(draws, d0):
    """ReferenceOrderNoise registered as a ``diffndm_b200.sampler.NoiseProvider`` (imports torch-side code lazily)."""
    from diffndm_b200.sampler import NoiseProvider
    cls = type('ReferenceOrderNoiseProvider', (ReferenceOrderNoise, NoiseProvider), {})
    return cls(draws, d0)


class ReferenceOrderNoise:
    """``diffndm_b200.sampler.NoiseProvider`` over the flat list of draws of a reference run: hands the sampler the draws of
    a whole event, reading them in the order the reference's sequential code made them (conditional_model.py:760-800 for an
    SPSA update, :1095-1128 for an ATP event, :1262-1306 for the s == 30 branch)."""

    def __init__(self, draws, d0):
        self.draws, self.i = draws, int(d0)

    def _next(self, shape):
        a = self.draws[self.i]
        assert tuple(a.shape) == tuple(shape), (self.i, a.shape, shape)
        self.i += 1
        return a

    def step(self, n_l):
        return self._next((n_l, self.draws[self.i].shape[1]))

    def spsa(self, k, sizes):
        n_l = int(sum(sizes))
        pert, plus, minus = [], [], []
        for _ in range(k):
            pert.append(np.concatenate([self._next((int(n), 3)) for n in sizes]))
            D = self.draws[self.i].shape[1]
            plus.append(self._next((n_l, D)))
            minus.append(self._next((n_l, D)))
        return np.stack(pert), np.stack(plus + minus)

    def atp(self, n_groups, n_l):
        D = self.draws[self.i].shape[1]
        x0 = [self._next((n_l, D))]
        steps = []
        for _ in range(n_groups - 1):
            steps.append(self._next((n_l, D)))
            x0.append(self._next((n_l, D)))
        return np.stack(steps), np.stack(x0)

    def mixed(self, n_groups, k, sizes):
        n_l = int(sum(sizes))
        D = self.draws[self.i].shape[1]
        x0 = [self._next((n_l, D))]
        steps, spsa = [], []
        for _ in range(n_groups - 1):
            steps.append(self._next((n_l, D)))
            spsa.append(self.spsa(k, sizes))
            x0.append(self._next((n_l, D)))
        return np.stack(steps), spsa, np.stack(x0)
