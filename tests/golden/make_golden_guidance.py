"""Generate ``guidance.npz``: the reference's OWN guidance code (build container only) on seeded inputs.

    python tests/golden/make_golden_guidance.py

Runs, unmodified, ``ConditionalDDPM.my_update_z_lig`` (SPSA, conditional_model.py:760-813), the ATP block of
``sample_given_pocket`` (:1085-1241), the ``s == 30`` mixed branch (:1261-1418) and the per-event feature rescaling
(:1235-1240, 1253-1258) inside a complete ``sample_given_pocket(svdd=1, spsa=1)`` with the real 500-step schedule.
Only the chemistry is stubbed (RDKit / OpenBabel are absent): ``handle_to_mol`` keeps its in-place translation and hands
(coordinates, atom types) to ``guidance_common.geometric_reward``, which stands in for ``my_reward_for_SPSA / _SVDD``.
Every Gaussian draw comes from numpy PCG64 through a patched ``torch.randn`` (``NoiseStream``), so the fixture stores the
seed and the draw shapes instead of the noise; the denoiser is wrapped in ``SyntheticScoreDynamics`` so that the states
stay O(1..10) A (random-init weights cannot denoise).  The reference hard-codes 20 samples per pocket in its re-batching
(``i * 20``, :1152-1175), hence B = 20.
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
from guidance_common import SyntheticScoreDynamics, geometric_reward, NoiseStream  # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init, weights_checksum  # noqa: E402
from diffndm_b200 import synthetic  # noqa: E402

T = torch.from_numpy
K_SPSA = 10            # hard-coded in my_update_z_lig (:764)


def npy(t):
    return t.detach().cpu().numpy().copy()


class Recorder:
    """Wraps the methods of one ConditionalDDPM instance and logs every call in order."""

    def __init__(self, ddpm, stream):
        self.ddpm, self.stream, self.log, self.depth = ddpm, stream, [], 0
        self.orig = {n: getattr(ddpm, n) for n in ('sample_p_zs_given_zt', 'my_to_x0', 'my_update_z_lig',
                                                    'sample_p_xh_given_z0')}
        ddpm.sample_p_zs_given_zt = self.step
        ddpm.my_to_x0 = self.to_x0
        ddpm.my_update_z_lig = self.update
        ddpm.sample_p_xh_given_z0 = self.xh0
        ddpm.handle_to_mol = self.handle_to_mol
        ddpm.my_reward_for_SPSA = self.reward
        ddpm.my_reward_for_SVDD = self.reward
        ddpm.my_reward_function = lambda *a, **k: 0.0

    def restore(self):
        for n, f in self.orig.items():
            setattr(self.ddpm, n, f)

    def _enter(self, kind, **kw):
        e = dict(kind=kind, depth=self.depth, d0=len(self.stream.shapes), **kw)
        self.log.append(e)
        self.depth += 1
        return e

    def _exit(self, e, **kw):
        self.depth -= 1
        e['d1'] = len(self.stream.shapes)
        e.update(kw)

    def step(self, s, t, z, xp, lm, pm, optimize, fix_noise=False):
        e = self._enter('step', s=npy(s), t=npy(t), z_in=npy(z), xp_in=npy(xp), lm=npy(lm))
        out = self.orig['sample_p_zs_given_zt'](s, t, z, xp, lm, pm, optimize, fix_noise)
        self._exit(e, z_out=npy(out[0]), xp_out=npy(out[1]))
        return out

    def to_x0(self, t, z, xp, lm, pm, n):
        e = self._enter('x0', t=npy(t), z_in=npy(z), xp_in=npy(xp), lm=npy(lm))
        out = self.orig['my_to_x0'](t, z, xp, lm, pm, n)
        self._exit(e, x=npy(out[0]), h=npy(out[1]).argmax(1).astype(np.int8))
        return out

    def xh0(self, z, xp, lm, pm, n, fix_noise=False):
        e = self._enter('xh0', z_in=npy(z), xp_in=npy(xp), lm=npy(lm))
        out = self.orig['sample_p_xh_given_z0'](z, xp, lm, pm, n, fix_noise)
        self._exit(e)
        return out

    def update(self, z, xp, lm, pm, com_before, dataset_info, sanitize, relax_iter, largest_frag, t_array, n_samples,
               zeta, guidance_scale=1e-2):
        e = self._enter('spsa', z_in=npy(z), xp_in=npy(xp), lm=npy(lm), t=npy(t_array), zeta=float(zeta),
                        guidance_scale=float(guidance_scale))
        out = self.orig['my_update_z_lig'](z, xp, lm, pm, com_before, dataset_info, sanitize, relax_iter, largest_frag,
                                           t_array, n_samples, zeta, guidance_scale=guidance_scale)
        self._exit(e, z_out=npy(out[0]), xp_out=npy(out[1]))
        return out

    def handle_to_mol(self, xh_lig, xh_pocket, lig_mask, pocket_mask, pocket_com_before, *a, **k):
        # the translation back to the input frame, IN PLACE like the reference (conditional_model.py:845-864)
        n = int(pocket_mask.max()) + 1
        cnt = torch.bincount(pocket_mask, minlength=n).clamp(min=1).to(xh_pocket.dtype)
        com_after = torch.zeros((n, 3), dtype=xh_pocket.dtype).index_add_(0, pocket_mask, xh_pocket[:, :3]) / cnt[:, None]
        xh_pocket[:, :3] += (pocket_com_before - com_after)[pocket_mask]
        xh_lig[:, :3] += (pocket_com_before - com_after)[lig_mask]
        mols = (npy(xh_lig[:, :3]), npy(xh_lig[:, 3:].argmax(1)), npy(lig_mask))
        self.log.append(dict(kind='mol', depth=self.depth, x=mols[0], types=mols[1].astype(np.int8), lm=mols[2]))
        return [mols]

    def reward(self, molecules):
        x, types, lm = molecules[0]
        r = geometric_reward(x, types, lm)
        self.log.append(dict(kind='reward', depth=self.depth, r=np.asarray(r, np.float64)))
        return r


def make_case(cfg, W, seed, B, n_atoms, n_pocket, timesteps=500):
    """Inputs of one run: synthetic pocket repeated B times, equally sized ligands, a data point x_0 in the cavity."""
    px, pt = synthetic.synthetic_pocket(100 + seed, n_pocket)
    sizes = np.full(B, n_atoms, np.int64)
    pose = synthetic.synthetic_ligand_pose(seed, sizes[:1], px.mean(0), r_max=3.0)           # one pose shared by all samples
    x0_rel = pose.copy()
    x0_rel[:, :3] -= px[0]                                                                  # relative to the pocket's first atom
    return px, pt, sizes, x0_rel


def run_full(cfg, W, seed, B, n_atoms, n_pocket, svdd, spsa, timesteps=500):
    px, pt, sizes, x0_rel = make_case(cfg, W, seed, B, n_atoms, n_pocket)
    dyn, ddpm = build_reference_model(cfg, W, timesteps=timesteps)
    ddpm.dynamics = SyntheticScoreDynamics(dyn, x0_rel)
    stream = NoiseStream(seed)
    rec = Recorder(ddpm, stream)
    onehot = np.eye(cfg.atom_nf, dtype=np.float32)[pt]
    pocket = {'x': T(np.tile(px, (B, 1))), 'one_hot': T(np.tile(onehot, (B, 1))),
              'size': torch.tensor([n_pocket] * B), 'mask': torch.arange(B).repeat_interleave(n_pocket)}
    com_before = T(np.tile(px.mean(0, dtype=np.float64).astype(np.float32), (B, 1)))
    keep = torch.randn
    torch.randn = stream.randn
    try:
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            xh_lig, xh_pocket, lig_mask, pocket_mask = ddpm.sample_given_pocket(
                pocket, torch.tensor(sizes), com_before, None, False, 0, False, 'x', 'cpu', 0, None, None, svdd, spsa,
                timesteps=timesteps)
    finally:
        torch.randn = keep
        rec.restore()
    inputs = dict(pocket_x=px, pocket_t=pt, sizes=sizes, x0_rel=x0_rel, com_before=npy(com_before))
    final = dict(final_lig=npy(xh_lig), final_pocket=npy(xh_pocket), final_lig_mask=npy(lig_mask))
    return inputs, final, rec.log, stream


def run_inpaint(cfg, W, seed, B, n_atoms, n_fixed, n_pocket, timesteps, resamplings, svdd=1):
    """The unmodified ``ConditionalDDPM.inpaint`` (conditional_model.py:1491-1790) with its hard-wired SPSA window
    (12 <= s <= 16, first resampling) and, with svdd = 1, the ATP block at s <= 10, s % 2 == 0 (:1629-1778)."""
    px, pt, sizes, x0_rel = make_case(cfg, W, seed, B, n_atoms, n_pocket)
    dyn, ddpm = build_reference_model(cfg, W)
    ddpm.dynamics = SyntheticScoreDynamics(dyn, x0_rel)
    stream = NoiseStream(seed)
    rec = Recorder(ddpm, stream)
    onehot = np.eye(cfg.atom_nf, dtype=np.float32)
    pocket = {'x': T(np.tile(px, (B, 1))), 'one_hot': T(np.tile(onehot[pt], (B, 1))),
              'size': torch.tensor([n_pocket] * B), 'mask': torch.arange(B).repeat_interleave(n_pocket)}
    lig_x = np.tile(x0_rel[:, :3] + px[0], (B, 1)).astype(np.float32)                       # the pose itself is the input ligand
    lig_t = np.tile(x0_rel[:, 3:].argmax(1), B)
    fixed = np.tile((np.arange(n_atoms) < n_fixed).astype(np.float32), B)
    lig_mask = np.repeat(np.arange(B), n_atoms)
    ligand = {'x': T(lig_x.copy()), 'one_hot': T(onehot[lig_t].copy()), 'size': torch.tensor(sizes), 'mask': T(lig_mask)}
    com_before = T(np.tile(px.mean(0, dtype=np.float64).astype(np.float32), (B, 1)))
    keep = torch.randn
    torch.randn = stream.randn
    try:
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            xh_lig, xh_pocket, lm, pm = ddpm.inpaint(ligand, pocket, T(fixed), svdd, com_before, None, False, 0, False,
                                                     resamplings=resamplings, return_frames=1, timesteps=timesteps, center='ligand')
    finally:
        torch.randn = keep
        rec.restore()
    inputs = dict(pocket_x=px, pocket_t=pt, sizes=sizes, x0_rel=x0_rel, com_before=npy(com_before), lig_x=lig_x, lig_t=lig_t,
                  lig_fixed=fixed, timesteps=np.asarray(timesteps), resamplings=np.asarray(resamplings))
    final = dict(final_lig=npy(xh_lig), final_pocket=npy(xh_pocket))
    return inputs, final, rec.log, stream


def parse_inpaint_log(log, timesteps, resamplings, svdd, G=5):
    it = iter(log)

    def nxt(kind, depth):
        e = next(it)
        assert e['kind'] == kind and e['depth'] == depth, (e['kind'], e['depth'], kind, depth)
        return e

    events = []
    for s in reversed(range(timesteps)):
        for u in range(resamplings):
            st = nxt('step', 0)
            events.append(dict(kind='step', s=s, u=u, d0=st['d0'], z_in=st['z_in'], xp_in=st['xp_in'], lm=st['lm']))
            if 12 <= s <= 16 and u < 1:
                up = nxt('spsa', 0)
                ev = dict(kind='spsa', s=s, d0=up['d0'], z_in=up['z_in'], xp_in=up['xp_in'], lm=up['lm'], zeta=up['zeta'],
                          guidance_scale=up['guidance_scale'], z_out=up['z_out'], xp_out=up['xp_out'])
                plus_x, minus_x, fp, fm = [], [], [], []
                for _ in range(K_SPSA):
                    nxt('x0', 1); nxt('xh0', 2); nxt('x0', 1); nxt('xh0', 2)
                    ma = nxt('mol', 1); mb = nxt('mol', 1); ra = nxt('reward', 1); rb = nxt('reward', 1)
                    plus_x.append(ma['x']); minus_x.append(mb['x']); fp.append(ra['r']); fm.append(rb['r'])
                ev.update(mol_x=np.stack(plus_x + minus_x).astype(np.float32), f_plus=np.stack(fp), f_minus=np.stack(fm))
                events.append(ev)
        if svdd == 1 and s <= 10 and s % 2 == 0:
            x0c = nxt('x0', 0); nxt('xh0', 1)
            ev = dict(kind='atp', s=s, d0=x0c['d0'], z_in=x0c['z_in'], xp0=x0c['xp_in'], lm=x0c['lm'])
            cand_z = []
            for i in range(G - 1):
                c = nxt('step', 0)
                if i == 0:
                    ev['xp_in'] = c['xp_in']                         # the translated pocket of the current state
                cand_z.append(c['z_out'])
                nxt('x0', 0); nxt('xh0', 1)
            nxt('mol', 0); r0 = nxt('reward', 0); nxt('mol', 0); r1 = nxt('reward', 0)
            ev.update(cand_z=np.stack(cand_z), r0=r0['r'], r=r1['r'])
            events.append(ev)
    fin = nxt('xh0', 0)
    events.append(dict(kind='final', d0=fin['d0'], z_in=fin['z_in'], xp_in=fin['xp_in'], lm=fin['lm']))
    for a, b in zip(events[:-1], events[1:]):
        if a['kind'] == 'atp':
            a['z_after'], a['xp_after'] = b['z_in'], b['xp_in']
    return events


def parse_log(log, timesteps, svdd, spsa, G=5):
    """Cut the flat call log into the events of the sampling loop (conditional_model.py:944-1420)."""
    it = iter([e for e in log])
    buf = []

    def nxt(kind=None, depth=None):
        e = buf.pop(0) if buf else next(it)
        assert kind is None or e['kind'] == kind, (e['kind'], kind)
        assert depth is None or e['depth'] == depth, (e['kind'], e['depth'], depth)
        return e

    def peek():
        if not buf:
            buf.append(next(it))
        return buf[0]

    def spsa_children(ev):
        """Entries logged inside one my_update_z_lig call: per i: x0(+), x0(-) [each with a nested xh0], mol(+), mol(-),
        reward(+), reward(-)."""
        plus_x, minus_x, plus_t, minus_t, fp, fm = [], [], [], [], [], []
        for _ in range(K_SPSA):
            a = nxt('x0', 1); nxt('xh0', 2)
            b = nxt('x0', 1); nxt('xh0', 2)
            ma = nxt('mol', 1); mb = nxt('mol', 1)
            ra = nxt('reward', 1); rb = nxt('reward', 1)
            plus_x.append(ma['x']); minus_x.append(mb['x']); plus_t.append(ma['types']); minus_t.append(mb['types'])
            fp.append(ra['r']); fm.append(rb['r'])
        ev.update(mol_x=np.stack(plus_x + minus_x).astype(np.float32), mol_t=np.stack(plus_t + minus_t),
                  f_plus=np.stack(fp), f_minus=np.stack(fm))

    events = []
    for s in reversed(range(timesteps)):
        st = nxt('step', 0)
        assert int(round(float(st['s'][0, 0]) * timesteps)) == s
        events.append(dict(kind='step', s=s, d0=st['d0'], z_in=st['z_in'], xp_in=st['xp_in'], lm=st['lm'],
                           z_out=st['z_out'], xp_out=st['xp_out']))
        if svdd == 1 and s <= 50 and s % 10 == 0:
            ev = dict(kind='atp', s=s, z_in=st['z_out'], xp_in=st['xp_out'], lm=st['lm'])
            x0c = nxt('x0', 0); nxt('xh0', 1)
            ev['d0'] = x0c['d0']
            cand_z = []
            for _ in range(G - 1):
                c = nxt('step', 0)
                cand_z.append(c['z_out'])
                nxt('x0', 0); nxt('xh0', 1)
            m0 = nxt('mol', 0); r0 = nxt('reward', 0); m1 = nxt('mol', 0); r1 = nxt('reward', 0)
            ev.update(cand_z=np.stack(cand_z), r0=r0['r'], r=r1['r'], mol0_x=m0['x'].astype(np.float32), mol0_t=m0['types'])
            events.append(ev)
        if spsa == 1 and s <= 30 and s % 2 == 0:
            u = nxt('spsa', 0)
            ev = dict(kind='spsa', s=s, d0=u['d0'], z_in=u['z_in'], xp_in=u['xp_in'], lm=u['lm'], zeta=u['zeta'],
                      guidance_scale=u['guidance_scale'], z_out=u['z_out'], xp_out=u['xp_out'])
            spsa_children(ev)
            events.append(ev)
            if s == 30:
                x0c = nxt('x0', 0); nxt('xh0', 1)
                ev = dict(kind='mixed', s=s, d0=x0c['d0'], z_in=x0c['z_in'], xp_in=x0c['xp_in'], lm=x0c['lm'])
                subs = []
                for i in range(G - 1):
                    c = nxt('step', 0)
                    u = nxt('spsa', 0)
                    sub = dict(step_z_in=c['z_in'], step_z_out=c['z_out'], zeta=u['zeta'], z_out=u['z_out'])
                    spsa_children(sub)
                    x0i = nxt('x0', 0); nxt('xh0', 1)
                    sub['cand_z'] = x0i['z_in']                     # the rescaled candidate that enters the selection
                    subs.append(sub)
                m0 = nxt('mol', 0); r0 = nxt('reward', 0); m1 = nxt('mol', 0); r1 = nxt('reward', 0)
                ev.update(r0=r0['r'], r=r1['r'])
                for i, sub in enumerate(subs):
                    for k, v in sub.items():
                        ev[f'sub{i}_{k}'] = np.asarray(v)
                events.append(ev)
    fin = nxt('xh0', 0)
    events.append(dict(kind='final', d0=fin['d0'], z_in=fin['z_in'], xp_in=fin['xp_in'], lm=fin['lm']))
    # the state an event leaves behind is the input of the next call
    for a, b in zip(events[:-1], events[1:]):
        if a['kind'] in ('atp', 'spsa', 'mixed'):
            a['z_after'], a['xp_after'], a['lm_after'] = b['z_in'], b['xp_in'], b['lm']
    return events


def main():
    torch.set_num_threads(8)
    cfg = DynamicsConfig()
    seed_w, gain = 0, 0.3
    W = random_init(cfg, seed_w, gain)
    out = dict(weight_seed=np.asarray(seed_w), coord_head_gain=np.asarray(gain), weights_checksum=np.asarray(weights_checksum(W)))

    KEEP_EVENTS = {('atp', 50), ('atp', 40), ('atp', 30), ('atp', 20), ('spsa', 30), ('spsa', 28), ('spsa', 0), ('mixed', 30)}
    KEEP_STEPS = {50, 45, 41, 39, 31, 29}                  # plain steps: input state only (trajectory.npz pins the step itself)

    def store(case, inputs, final, events, stream):
        for k, v in {**inputs, **final}.items():
            out[f'{case}/{k}'] = np.asarray(v)
        out[f'{case}/noise_seed'] = np.asarray(stream_seed[case])
        out[f'{case}/draw_shapes'] = np.asarray([list(s) + [0] * (2 - len(s)) for s in stream.shapes], np.int16)
        index = []                                            # (kind, s, first draw) of EVERY event, in loop order
        for i, ev in enumerate(events):
            index.append((['step', 'atp', 'spsa', 'mixed', 'final'].index(ev['kind']), ev.get('s', -1), ev['d0']))
            if ev['kind'] == 'step':
                if ev['s'] in KEEP_STEPS:
                    out[f'{case}/step{ev["s"]}/z_in'] = ev['z_in']
                    if ev['s'] == max(KEEP_STEPS):
                        out[f'{case}/step{ev["s"]}/xp_in'] = ev['xp_in']
                        out[f'{case}/step{ev["s"]}/lm'] = ev['lm']
                continue
            if ev['kind'] == 'final':
                out[f'{case}/final/z_in'] = ev['z_in']
                continue
            if (ev['kind'], ev['s']) not in KEEP_EVENTS:
                continue
            for k, v in ev.items():
                if k not in ('kind', 's', 'd0'):
                    out[f'{case}/{ev["kind"]}{ev["s"]}/{k}'] = np.asarray(v)
        out[f'{case}/event_index'] = np.asarray(index, np.int32)

    stream_seed = {}
    # the BASELINE "mixed SPSA+ATP" configuration on a small pocket: every event type, real 500-step schedule
    case = 'mixed_b20'
    stream_seed[case] = 2024
    inputs, final, log, stream = run_full(cfg, W, stream_seed[case], B=20, n_atoms=3, n_pocket=8, svdd=1, spsa=1)
    events = parse_log(log, 500, 1, 1)
    store(case, inputs, final, events, stream)
    print(case, 'events', {k: sum(e['kind'] == k for e in events) for k in ('step', 'atp', 'spsa', 'mixed')},
          'draws', len(stream.shapes), '|z| final', np.abs(final['final_lig'][:, :3]).max())

    # inpainting with its guidance branches: SPSA window 12 <= s <= 16, ATP at s <= 10 (svdd = 1)
    case = 'inpaint_b20'
    stream_seed[case] = 2025
    inputs, final, log, stream = run_inpaint(cfg, W, stream_seed[case], B=20, n_atoms=4, n_fixed=2, n_pocket=8, timesteps=20,
                                             resamplings=2, svdd=1)
    ev_i = parse_inpaint_log(log, 20, 2, 1)
    for k, v in {**inputs, **final}.items():
        out[f'{case}/{k}'] = np.asarray(v)
    out[f'{case}/noise_seed'] = np.asarray(stream_seed[case])
    out[f'{case}/draw_shapes'] = np.asarray([list(sh) + [0] * (2 - len(sh)) for sh in stream.shapes], np.int16)
    index = []
    for ev in ev_i:
        index.append((['step', 'atp', 'spsa', 'mixed', 'final'].index(ev['kind']), ev.get('s', -1), ev['d0']))
        if (ev['kind'], ev.get('s')) in {('spsa', 16), ('spsa', 15), ('atp', 10), ('atp', 8)}:
            for k, v in ev.items():
                if k not in ('kind', 's', 'd0'):
                    out[f'{case}/{ev["kind"]}{ev["s"]}/{k}'] = np.asarray(v)
    out[f'{case}/event_index'] = np.asarray(index, np.int32)
    print(case, 'events', {k: sum(e['kind'] == k for e in ev_i) for k in ('step', 'atp', 'spsa')}, 'draws', len(stream.shapes),
          '|x| final', np.abs(final['final_lig'][:, :3]).max())

    path = os.path.join(HERE, 'guidance.npz')
    np.savez_compressed(path, **out)
    print('guidance.npz', os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
