"""Generate ``inpaint.npz`` FROM THE REFERENCE ITSELF (build container only).

    python tests/golden/make_golden_inpaint.py

Runs the unmodified ``ConditionalDDPM.inpaint`` (conditional_model.py:1491-1790, RePaint-style resampling) on seeded
inputs with every Gaussian draw recorded, and stores the draws, the state after every (s, u) resampling iteration and
the final molecules.  ``timesteps`` stays below 12 so that the reference's hard-wired SPSA window (12 <= s <= 16, which
needs the host chemistry) is not entered.  Weights are regenerated from ``random_init(seed)`` and guarded by a checksum.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_loader import build_reference_model  # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init, weights_checksum  # noqa: E402
from diffndm_b200 import synthetic  # noqa: E402

T = torch.from_numpy


def main():
    torch.set_num_threads(8)
    cfg = DynamicsConfig()
    seed_w, gain = 0, 0.3
    W = random_init(cfg, seed_w, gain)
    _, ddpm = build_reference_model(cfg, W)
    meta = dict(weight_seed=seed_w, coord_head_gain=gain, weights_checksum=weights_checksum(W))

    noises = []

    def rec_gauss(size, device):
        x = torch.randn(size, device=device)
        noises.append(x.numpy().copy())
        return x

    ddpm.sample_gaussian = rec_gauss

    def case(name, px, pt, sizes, n_fixed, seed, timesteps, resamplings, center='ligand'):
        noises.clear()
        torch.manual_seed(seed)
        rng = np.random.default_rng(seed)
        B, n_p = len(sizes), len(px)
        onehot = np.eye(cfg.atom_nf, dtype=np.float32)
        pocket = {'x': T(np.tile(px, (B, 1))), 'one_hot': T(np.tile(onehot[pt], (B, 1))),
                  'size': torch.tensor([n_p] * B), 'mask': torch.arange(B).repeat_interleave(n_p)}
        lig_mask = np.repeat(np.arange(B), sizes)
        # input ligand: a compact cloud near the pocket centre; the first n_fixed atoms of every ligand are kept
        lx = (px.mean(0)[None] + rng.normal(size=(len(lig_mask), 3)) * 1.5).astype(np.float32)
        lt = rng.integers(0, cfg.atom_nf, size=len(lig_mask))
        fixed = np.concatenate([(np.arange(n) < n_fixed) for n in sizes]).astype(np.float32)
        ligand = {'x': T(lx.copy()), 'one_hot': T(onehot[lt].copy()), 'size': torch.tensor(sizes), 'mask': T(lig_mask)}
        states = []
        orig_combine_probe = ddpm.sample_p_xh_given_z0

        def rec_final(z0, xp, lm, pm, n):
            states.append(dict(z_final_in=z0.detach().numpy().copy(), xp_final_in=xp.detach().numpy().copy()))
            return orig_combine_probe(z0, xp, lm, pm, n)
        ddpm.sample_p_xh_given_z0 = rec_final
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            out_lig, out_pocket, lm, pm = ddpm.inpaint(
                ligand, pocket, T(fixed), 0, torch.zeros(B, 3), None, False, 0, False, resamplings=resamplings,
                return_frames=1, timesteps=timesteps, center=center)
        ddpm.sample_p_xh_given_z0 = orig_combine_probe
        # draws: 1 (z_T) + per (s,u): p(z_s|z_t), q(z_s|x) [+ q(z_t|z_s) if u < R-1] + 1 (p(x|z_0))
        expect = 1 + timesteps * (2 * resamplings + (resamplings - 1)) + 1
        assert len(noises) == expect, (len(noises), expect)
        out = dict(pocket_x=px, pocket_t=pt, sizes=np.asarray(sizes), timesteps=np.int64(timesteps),
                   resamplings=np.int64(resamplings), lig_x=lx, lig_t=lt, lig_fixed=fixed, lig_mask=lig_mask,
                   noise=np.stack(noises), final_lig=out_lig.numpy(), final_pocket=out_pocket.numpy(),
                   z_final_in=states[0]['z_final_in'], xp_final_in=states[0]['xp_final_in'])
        print(f'inpaint[{name}]: T={timesteps} R={resamplings} N_l={len(lig_mask)} draws={len(noises)} '
              f'|x|max={np.abs(out_lig.numpy()[:, :3]).max():.3f}')
        return {f'{name}/{k}': v for k, v in out.items()}

    fx = {}
    sx, st = synthetic.synthetic_pocket(41, 48)
    fx.update(case('synth48_b3_T6_R2', sx, st, [9, 6, 12], 4, 200, 6, 2))
    fx.update(case('synth48_b2_T4_R3', sx, st, [7, 10], 3, 201, 4, 3))
    for k, v in meta.items():
        fx[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, 'inpaint.npz'), **fx)
    print('inpaint.npz', os.path.getsize(os.path.join(HERE, 'inpaint.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
