"""Generate ``parity_r2.npz`` FROM THE REFERENCE ITSELF (build container only): the round-2 parity cases.

    python tests/golden/make_golden_r2.py

* trajectories with realistic states: the unmodified ``sample_given_pocket`` over a ``SyntheticScoreDynamics``-wrapped
  denoiser (guidance_common.py), so |z| stays O(1..10) A and the per-step coordinate bar of 1e-3 A is asserted in A --
  3rfm (286 atoms, real PDB coordinates) with the full 500-step schedule and a 600-atom synthetic pocket;
* forward cases: untied ``coord_mlp.4`` / ``cross_product_mlp.4`` (a loaded checkpoint carries both keys), a radial stress
  case (r^2-column weights x 20, diffuse ligands with r^2 up to several hundred), the 5ndu pocket, a 600-atom pocket;
* ``ConditionalDDPM.inpaint`` on the 5ndu pocket with the 17 atoms of example/fragments.sdf fixed and 10 new atoms, and on a
  600-atom pocket (BASELINE configs[4]).

Gaussian draws come from numpy PCG64 through a patched ``torch.randn``: the fixture stores seeds and shapes, not noise.
Weights are regenerated from ``random_init(seed, ...)`` and guarded by checksums.
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
from guidance_common import SyntheticScoreDynamics, NoiseStream, stress_weights  # noqa: E402
from make_golden import pocket_3rfm, pocket_5ndu  # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init, weights_checksum  # noqa: E402
from diffndm_b200 import synthetic  # noqa: E402

T = torch.from_numpy
npy = lambda t: t.detach().cpu().numpy().copy()


def pose_rel(seed, sizes, px):
    pose = synthetic.synthetic_ligand_pose(seed, np.asarray(sizes), px.mean(0))
    pose[:, :3] -= px[0]
    return pose


def read_sdf_atoms(path):
    with open(path) as f:
        lines = f.read().splitlines()
    n = int(lines[3][:3])
    xyz = np.array([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in lines[4:4 + n]], np.float32)
    el = [l[31:34].strip() for l in lines[4:4 + n]]
    return xyz, np.array([synthetic.ATOM_TYPES.index(e) for e in el], np.int64)


def main():
    torch.set_num_threads(8)
    cfg = DynamicsConfig()
    seed_w, gain = 0, 0.3
    W = random_init(cfg, seed_w, gain)
    out = dict(weight_seed=np.asarray(seed_w), coord_head_gain=np.asarray(gain), weights_checksum=np.asarray(weights_checksum(W)))
    only = os.environ.get('R2_ONLY')                     # regenerate one group, keep the rest of the existing file
    path = os.path.join(HERE, 'parity_r2.npz')
    if only:
        old = np.load(path)
        out.update({k: old[k] for k in old.files if not k.startswith(only)})
    want = lambda group: (not only) or group.startswith(only) or only.startswith(group)
    px3, pt3 = pocket_3rfm()
    px5, pt5 = pocket_5ndu()
    px6, pt6 = synthetic.synthetic_pocket(61, 600)
    out['pocket600_seed'] = np.asarray(61)

    # ------------------------------------------------------------------ trajectories --------------------------------------
    def traj_case(name, px, pt, sizes, seed, timesteps, keep):
        if not want('traj_' + name):
            return
        B, n_p = len(sizes), len(px)
        x0_rel = pose_rel(seed, sizes, px)
        dyn, ddpm = build_reference_model(cfg, W)
        ddpm.dynamics = SyntheticScoreDynamics(dyn, x0_rel)
        ddpm.handle_to_mol = lambda *a, **k: [[]]
        ddpm.my_reward_function = lambda *a, **k: 0.0
        stream = NoiseStream(seed)
        states = {}
        orig = ddpm.sample_p_zs_given_zt

        def rec(s, t, z, xp, lm, pm, optimize, fix_noise=False):
            d0 = len(stream.shapes)
            zi, xpi = npy(z), npy(xp)
            o = orig(s, t, z, xp, lm, pm, optimize, fix_noise)
            si = int(round(float(s[0, 0]) * timesteps))
            if si in keep:
                states[si] = dict(d0=np.asarray(d0), z_in=zi, xp_in=xpi, z_out=npy(o[0]), xp_out=npy(o[1]))
            return o
        ddpm.sample_p_zs_given_zt = rec
        onehot = np.eye(cfg.atom_nf, dtype=np.float32)[pt]
        pocket = {'x': T(np.tile(px, (B, 1))), 'one_hot': T(np.tile(onehot, (B, 1))),
                  'size': torch.tensor([n_p] * B), 'mask': torch.arange(B).repeat_interleave(n_p)}
        keep_randn, torch.randn = torch.randn, stream.randn
        try:
            with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
                xh_lig, xh_pocket, lm, pm = ddpm.sample_given_pocket(
                    pocket, torch.tensor(sizes), torch.zeros(B, 3), None, False, 0, False, 'x', 'cpu', 0, None, None, 0, 0,
                    timesteps=timesteps)
        finally:
            torch.randn = keep_randn
        assert len(stream.shapes) == timesteps + 2
        o = dict(pocket_x=px, pocket_t=pt, sizes=np.asarray(sizes), timesteps=np.asarray(timesteps), x0_rel=x0_rel,
                 noise_seed=np.asarray(seed), final_lig=npy(xh_lig), final_pocket=npy(xh_pocket))
        zmax = 0.0
        for si, st in states.items():
            zmax = max(zmax, np.abs(st['z_in'][:, :3]).max())
            for k, v in st.items():
                o[f's{si}_{k}'] = v
        o['kept_steps'] = np.asarray(sorted(states))
        print(f'traj[{name}]: T={timesteps} N_l={sum(sizes)} N_p={B * n_p} kept={sorted(states)} |z_x|max={zmax:.2f} '
              f'final |x|max={np.abs(npy(xh_lig)[:, :3]).max():.2f}')
        for k, v in o.items():
            out[f'traj_{name}/{k}'] = v

    traj_case('3rfm_b2', px3, pt3, [14, 20], 301, 500, {499, 450, 350, 250, 150, 80, 40, 20, 10, 5, 1, 0})
    traj_case('synth600_b2', px6, pt6, [23, 31], 302, 50, {49, 40, 25, 10, 3, 0})

    # ------------------------------------------------------------------ forward cases -------------------------------------
    def forward_case(name, Wc, px, pt, sizes, seed, t_vals, spread=1.0, extra=None, untie=False):
        if not want('fwd_' + name):
            return
        dyn64, _ = build_reference_model(cfg, Wc, dtype=torch.float64, untie_heads=untie)
        b = synthetic.make_batch(px, pt, np.asarray(sizes), seed)
        if spread != 1.0:
            for s in range(len(sizes)):
                m = b['lig_mask'] == s
                c = b['xh_lig'][m, :3].mean(0)
                b['xh_lig'][m, :3] = (b['xh_lig'][m, :3] - c) * spread + c
        t = np.asarray(t_vals, np.float32).reshape(-1, 1)
        trace = {}
        blk = dyn64.egnn._modules[f'e_block_{cfg.n_layers - 1}']
        hk = blk.register_forward_hook(lambda m, i, o: trace.__setitem__('h', npy(o[0])))
        with torch.no_grad():
            o64 = dyn64(T(b['xh_lig']).double(), T(b['xh_pocket']).double(), T(t).double(), T(b['lig_mask']), T(b['pocket_mask']))
        hk.remove()
        n_l = len(b['lig_mask'])
        o = dict(xh_lig=b['xh_lig'], xh_pocket=b['xh_pocket'], lig_mask=b['lig_mask'], pocket_mask=b['pocket_mask'], t=t,
                 out_lig_f64=npy(o64[0]), h_lig_last=trace['h'][:n_l].astype(np.float32), weights_checksum=np.asarray(weights_checksum(Wc)))
        o.update(extra or {})
        xl = b['xh_lig'][:, :3]
        r2max = max(((xl[b['lig_mask'] == s][:, None] - xl[b['lig_mask'] == s][None]) ** 2).sum(-1).max() for s in range(len(sizes)))
        print(f'forward[{name}]: N_l={n_l} N_p={len(b["pocket_mask"])} |eps_x|max={np.abs(o["out_lig_f64"][:, :3]).max():.3f} '
              f'|eps_h|max={np.abs(o["out_lig_f64"][:, 3:]).max():.3f} ll r2max={r2max:.0f}')
        for k, v in o.items():
            out[f'fwd_{name}/{k}'] = v

    sx, st = synthetic.synthetic_pocket(21, 60)
    forward_case('untied_synth60_b3', random_init(cfg, seed_w, gain, untie_heads=True), sx, st, [7, 9, 5], 10, [0.5, 0.3, 0.9],
                 spread=2.0, extra=dict(pocket_seed=np.asarray(21), pocket_n=np.asarray(60)), untie=True)
    forward_case('r2stress_3rfm_b2', stress_weights(W), px3, pt3, [23, 17], 12, [0.9, 0.2], spread=4.0)
    forward_case('5ndu_b2', W, px5, pt5, [27, 17], 14, [0.4, 0.05], spread=2.5)
    forward_case('synth600_b2', W, px6, pt6, [23, 50], 15, [0.6, 0.1], spread=2.5)

    # ------------------------------------------------------------------ inpainting ----------------------------------------
    def inpaint_case(name, px, pt, lig_x, lig_t, fixed, sizes, seed, timesteps, resamplings):
        if not want('inp_' + name):
            return
        B, n_p = len(sizes), len(px)
        lig_mask = np.repeat(np.arange(B), sizes)
        x0_rel = np.concatenate([lig_x - px[0], np.eye(cfg.atom_nf, dtype=np.float32)[lig_t] / 4.0], 1).astype(np.float32)
        dyn, ddpm = build_reference_model(cfg, W)
        ddpm.dynamics = SyntheticScoreDynamics(dyn, x0_rel)
        stream = NoiseStream(seed)
        onehot = np.eye(cfg.atom_nf, dtype=np.float32)
        pocket = {'x': T(np.tile(px, (B, 1))), 'one_hot': T(np.tile(onehot[pt], (B, 1))),
                  'size': torch.tensor([n_p] * B), 'mask': torch.arange(B).repeat_interleave(n_p)}
        ligand = {'x': T(lig_x.copy()), 'one_hot': T(onehot[lig_t].copy()), 'size': torch.tensor(sizes), 'mask': T(lig_mask)}
        keep_randn, torch.randn = torch.randn, stream.randn
        try:
            with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
                out_lig, out_pocket, lm, pm = ddpm.inpaint(
                    ligand, pocket, T(fixed), 0, torch.zeros(B, 3), None, False, 0, False, resamplings=resamplings,
                    return_frames=1, timesteps=timesteps, center='ligand')
        finally:
            torch.randn = keep_randn
        expect = 1 + timesteps * (2 * resamplings + (resamplings - 1)) + 1
        assert len(stream.shapes) == expect, (len(stream.shapes), expect)
        o = dict(pocket_x=px, pocket_t=pt, sizes=np.asarray(sizes), timesteps=np.asarray(timesteps), resamplings=np.asarray(resamplings),
                 lig_x=lig_x, lig_t=lig_t, lig_fixed=fixed, lig_mask=lig_mask, x0_rel=x0_rel, noise_seed=np.asarray(seed),
                 n_draws=np.asarray(expect), final_lig=npy(out_lig), final_pocket=npy(out_pocket))
        print(f'inpaint[{name}]: T={timesteps} R={resamplings} N_l={len(lig_mask)} N_p={B * n_p} '
              f'|x|max={np.abs(npy(out_lig)[:, :3]).max():.2f} fixed-atom drift={np.abs(npy(out_lig)[fixed > 0][:, :3] - (lig_x[fixed > 0] - lig_x[fixed > 0].mean(0))).max():.2f}')
        for k, v in o.items():
            out[f'inp_{name}/{k}'] = v

    # 5ndu: the two fragments of example/fragments.sdf fixed (17 atoms), 10 new atoms between them, two samples
    fx_x, fx_t = read_sdf_atoms(os.path.join(REFERENCE_ROOT, 'example', 'fragments.sdf'))
    rng = np.random.default_rng(5)
    new_x = (fx_x.mean(0)[None] + rng.normal(size=(10, 3)) * 2.0).astype(np.float32)
    new_t = rng.integers(0, 3, size=10)
    one_x, one_t = np.concatenate([fx_x, new_x]), np.concatenate([fx_t, new_t])
    one_f = np.concatenate([np.ones(len(fx_x)), np.zeros(10)]).astype(np.float32)
    inpaint_case('5ndu_frag17_new10_b2', px5, pt5, np.tile(one_x, (2, 1)), np.tile(one_t, 2), np.tile(one_f, 2), [27, 27], 401, 6, 2)
    # 600-atom pocket, ragged ligands, first 6 atoms fixed
    sizes = [18, 25]
    pose = synthetic.synthetic_ligand_pose(402, np.asarray(sizes), px6.mean(0))
    lt = pose[:, 3:].argmax(1)
    fixed = np.concatenate([(np.arange(n) < 6) for n in sizes]).astype(np.float32)
    inpaint_case('synth600_b2', px6, pt6, pose[:, :3].copy(), lt, fixed, sizes, 402, 5, 2)

    np.savez_compressed(path, **out)
    print('parity_r2.npz', os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
