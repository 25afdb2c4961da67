"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py

imports ``/root/reference`` unmodified (behind ``ref_loader`` shims), runs its own
``EGNNDynamics.get_edges`` / ``EGNNDynamics.forward`` / ``ConditionalDDPM.sample_p_zs_given_zt`` /
``sample_given_pocket`` on seeded inputs, and writes small ``.npz`` files.  Weights are never stored:
they are regenerated from ``diffndm_b200.weights.random_init(seed)`` and guarded by a checksum.
The reference ships no golden vectors for this path, so these fixtures are the pin
for ``oracle/egnn_oracle.py`` and, through it, for the CUDA engine.
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

from ref_loader import build_reference_model, REFERENCE_ROOT  # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init, weights_checksum  # noqa: E402
from diffndm_b200 import synthetic  # noqa: E402

T = torch.from_numpy


def ref_edges(dyn, b):
    return dyn.get_edges(T(b['lig_mask']), T(b['pocket_mask']), T(b['xh_lig'][:, :3].copy()),
                         T(b['xh_pocket'][:, :3].copy())).numpy()


def boundary_pairs(b, cutoff=5.0, band=1e-4):
    """Count same-sample (pocket,pocket)/(ligand,pocket) pairs whose fp64 distance lies within
    ``band`` of the cutoff: these are the only pairs on which cdist and direct-difference
    arithmetic may legitimately disagree."""
    xl = b['xh_lig'][:, :3].astype(np.float64)
    xp = b['xh_pocket'][:, :3].astype(np.float64)
    n = 0
    for s in np.unique(b['pocket_mask']):
        p = xp[b['pocket_mask'] == s]
        l = xl[b['lig_mask'] == s]
        dpp = np.sqrt(((p[:, None] - p[None]) ** 2).sum(-1))
        dlp = np.sqrt(((l[:, None] - p[None]) ** 2).sum(-1))
        n += int((np.abs(dpp - cutoff) < band).sum()) + 2 * int((np.abs(dlp - cutoff) < band).sum())
    return n


def pocket_3rfm():
    lig = synthetic.read_sdf_coords(os.path.join(REFERENCE_ROOT, 'example', '3rfm_B_CFF.sdf'))
    px, pt = synthetic.read_pocket_from_pdb(os.path.join(REFERENCE_ROOT, 'example', '3rfm.pdb'), lig)
    return px, pt


def pocket_5ndu():
    lig = synthetic.read_sdf_coords(os.path.join(REFERENCE_ROOT, 'example', '5ndu_C_8V2.sdf'))
    px, pt = synthetic.read_pocket_from_pdb(os.path.join(REFERENCE_ROOT, 'example', '5ndu.pdb'), lig)
    return px, pt


def main():
    torch.set_num_threads(8)
    cfg = DynamicsConfig()
    seed_w, gain = 0, 0.3
    W = random_init(cfg, seed_w, gain)
    wsum = weights_checksum(W)
    dyn, ddpm = build_reference_model(cfg, W)
    dyn64, _ = build_reference_model(cfg, W, dtype=torch.float64)
    meta = dict(weight_seed=seed_w, coord_head_gain=gain, weights_checksum=wsum)

    # ---- real pockets travel as fixtures (also the C1 bench pocket) -------------------------
    px3, pt3 = pocket_3rfm()
    px5, pt5 = pocket_5ndu()
    np.savez_compressed(os.path.join(HERE, 'pockets.npz'), x_3rfm=px3, t_3rfm=pt3, x_5ndu=px5, t_5ndu=pt5)
    print('3rfm pocket', px3.shape, '5ndu pocket', px5.shape)

    # ---- (i) edge KATs ---------------------------------------------------------------------
    cases = {}
    rng = np.random.default_rng(7)

    def add_edge_case(name, px, pt, sizes, seed, spread=1.0, shift=None):
        b = synthetic.make_batch(px, pt, np.asarray(sizes), seed)
        if spread != 1.0:   # emulate late-trajectory (compact) or early (diffuse) ligands
            for s in range(len(sizes)):
                m = b['lig_mask'] == s
                c = b['xh_lig'][m, :3].mean(0)
                b['xh_lig'][m, :3] = (b['xh_lig'][m, :3] - c) * spread + c
        if shift is not None:
            b['xh_lig'][:, :3] += shift
            b['xh_pocket'][:, :3] += shift
        e = ref_edges(dyn, b)
        cases[name] = dict(xh_lig=b['xh_lig'], xh_pocket=b['xh_pocket'], lig_mask=b['lig_mask'],
                           pocket_mask=b['pocket_mask'], edges=e.astype(np.int32),
                           n_boundary=np.int64(boundary_pairs(b)))
        print(f'edges[{name}]: N_l={len(b["lig_mask"])} N_p={len(b["pocket_mask"])} E={e.shape[1]} '
              f'boundary_pairs={cases[name]["n_boundary"]}')

    add_edge_case('3rfm_b3', px3, pt3, [23, 14, 31], 1)
    add_edge_case('3rfm_b3_compact', px3, pt3, [23, 14, 31], 2, spread=2.5)
    add_edge_case('3rfm_b2_shift', px3, pt3, [5, 50], 3, spread=3.0, shift=np.float32([0.37, -1.21, 2.03]))
    add_edge_case('5ndu_b2', px5, pt5, [17, 27], 4, spread=2.0)
    sx, st = synthetic.synthetic_pocket(11, 150)
    add_edge_case('synth150_b4_ragged', sx, st, [1, 2, 50, 9], 5, spread=2.0)
    sx, st = synthetic.synthetic_pocket(12, 40)
    add_edge_case('synth40_b1', sx, st, [6], 6)
    flat = {}
    for k, v in cases.items():
        for kk, vv in v.items():
            flat[f'{k}/{kk}'] = vv
    np.savez_compressed(os.path.join(HERE, 'edges.npz'), **flat)

    # ---- (ii) forward KATs -----------------------------------------------------------------
    def forward_case(name, px, pt, sizes, seed, t_vals, spread=1.0, keep_rows=6):
        b = synthetic.make_batch(px, pt, np.asarray(sizes), seed)
        if spread != 1.0:
            for s in range(len(sizes)):
                m = b['lig_mask'] == s
                c = b['xh_lig'][m, :3].mean(0)
                b['xh_lig'][m, :3] = (b['xh_lig'][m, :3] - c) * spread + c
        t = np.asarray(t_vals, np.float32).reshape(-1, 1)
        trace = {}
        hooks = []
        for i in range(cfg.n_layers):
            blk = dyn64.egnn._modules[f'e_block_{i}']
            hooks.append(blk.register_forward_hook(
                lambda m, inp, out, i=i: trace.__setitem__(i, (out[0].detach().numpy().copy(),
                                                                out[1].detach().numpy().copy()))))
        with torch.no_grad():
            o32 = dyn(T(b['xh_lig']), T(b['xh_pocket']), T(t), T(b['lig_mask']), T(b['pocket_mask']))
            o64 = dyn64(T(b['xh_lig']).double(), T(b['xh_pocket']).double(), T(t).double(),
                        T(b['lig_mask']), T(b['pocket_mask']))
        for h in hooks:
            h.remove()
        n_l = len(b['lig_mask'])
        rows = np.concatenate([np.arange(min(keep_rows, n_l)), n_l + np.arange(keep_rows)])
        out = dict(xh_lig=b['xh_lig'], xh_pocket=b['xh_pocket'], lig_mask=b['lig_mask'],
                   pocket_mask=b['pocket_mask'], t=t,
                   out_lig_f32=o32[0].numpy(), out_pocket_f32=o32[1].numpy(),
                   out_lig_f64=o64[0].numpy(), out_pocket_f64=o64[1].numpy(), trace_rows=rows)
        for i in range(cfg.n_layers):
            out[f'h_rows_{i}'] = trace[i][0][rows]
            out[f'x_lig_{i}'] = trace[i][1][:n_l]
        d = np.abs(o32[0].numpy() - o64[0].numpy())
        print(f'forward[{name}]: N_l={n_l} N_p={len(b["pocket_mask"])} |eps_x|max={np.abs(o64[0].numpy()[:, :3]).max():.3f} '
              f'|eps_h|max={np.abs(o64[0].numpy()[:, 3:]).max():.3f} f32-vs-f64 x={d[:, :3].max():.2e} h={d[:, 3:].max():.2e}')
        return {f'{name}/{k}': v for k, v in out.items()}

    fw = {}
    sx, st = synthetic.synthetic_pocket(21, 60)
    fw.update(forward_case('synth60_b3', sx, st, [7, 9, 5], 10, [0.5, 0.5, 0.5], spread=2.0))
    fw.update(forward_case('synth60_b3_tmix', sx, st, [12, 1, 20], 11, [0.998, 0.402, 0.0], spread=1.5))
    fw.update(forward_case('3rfm_b2', px3, pt3, [23, 17], 12, [0.2, 0.2], spread=2.5))
    fw.update(forward_case('3rfm_b1_diffuse', px3, pt3, [30], 13, [1.0], spread=1.0))
    for k, v in meta.items():
        fw[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, 'forward.npz'), **fw)

    # ---- (iii) step KAT + (iv) short teacher-forced trajectory -------------------------------
    noises = []
    orig_gauss = type(ddpm).sample_gaussian

    def rec_gauss(size, device):
        x = torch.randn(size, device=device)
        noises.append(x.numpy().copy())
        return x

    ddpm.sample_gaussian = rec_gauss
    ddpm.handle_to_mol = lambda *a, **k: [[]]
    ddpm.my_reward_function = lambda *a, **k: 0.0

    def traj_case(name, px, pt, sizes, seed, timesteps):
        noises.clear()
        torch.manual_seed(seed)
        B = len(sizes)
        n_p = len(px)
        onehot = np.eye(cfg.atom_nf, dtype=np.float32)[pt]
        pocket = {'x': T(np.tile(px, (B, 1))), 'one_hot': T(np.tile(onehot, (B, 1))),
                  'size': torch.tensor([n_p] * B), 'mask': torch.arange(B).repeat_interleave(n_p)}
        states = []
        orig_step = ddpm.sample_p_zs_given_zt

        def rec_step(s, t, z, xp, lm, pm, optimize, fix_noise=False):
            zi, xpi = z.detach().numpy().copy(), xp.detach().numpy().copy()
            out = orig_step(s, t, z, xp, lm, pm, optimize, fix_noise)
            states.append(dict(s=s.numpy().copy(), t=t.numpy().copy(), z_in=zi, xp_in=xpi,
                               z_out=out[0].detach().numpy().copy(), xp_out=out[1].detach().numpy().copy()))
            return out
        ddpm.sample_p_zs_given_zt = rec_step
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            xh_lig, xh_pocket, lig_mask, pocket_mask = ddpm.sample_given_pocket(
                pocket, torch.tensor(sizes), torch.zeros(B, 3), None, False, 0, False, 'x', 'cpu',
                0, None, None, 0, 0, timesteps=timesteps)
        ddpm.sample_p_zs_given_zt = orig_step
        out = dict(pocket_x=px, pocket_t=pt, sizes=np.asarray(sizes), timesteps=np.int64(timesteps),
                   final_lig=xh_lig.numpy(), final_pocket=xh_pocket.numpy(),
                   lig_mask=lig_mask.numpy(), pocket_mask=pocket_mask.numpy())
        # noise draws: [0] = z_T draw, [1..timesteps] = per-step draws, [timesteps+1] = p(x|z0) draw
        assert len(noises) == timesteps + 2, len(noises)
        out['noise'] = np.stack(noises)
        for i, st_ in enumerate(states):
            for k, v in st_.items():
                out[f'step{i}/{k}'] = v
        print(f'traj[{name}]: steps={timesteps} N_l={len(lig_mask)} final |x|max={np.abs(xh_lig.numpy()[:, :3]).max():.3f} '
              f'types={np.bincount(xh_lig.numpy()[:, 3:].argmax(1), minlength=10).tolist()}')
        return {f'{name}/{k}': v for k, v in out.items()}

    tr = {}
    sx, st = synthetic.synthetic_pocket(31, 50)
    tr.update(traj_case('synth50_b3_T10', sx, st, [8, 5, 11], 100, 10))
    tr.update(traj_case('3rfm_b2_T5', px3, pt3, [14, 20], 101, 5))
    for k, v in meta.items():
        tr[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, 'trajectory.npz'), **tr)
    for f in ['pockets.npz', 'edges.npz', 'forward.npz', 'trajectory.npz']:
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
