"""A whole sampling job: many pockets x B ligands each, every rank pulling pockets from the shared ``PocketQueue`` --
the layout of the reference's ``my_test.py:68-90`` (one process per pocket, one ``<pocket>.sdf`` each) on one process per GPU.

Per pocket: build the pocket batch, sample (CUDA-graph replay of the reverse step, captured once per batch shape), perceive
bonds on the GPU, assemble molecules, keep the largest fragment, write one SDF file.  Used by ``scripts/sample_pockets.py``
and by the ``job`` section of ``bench.py``."""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import output, synthetic
from .parallel import PocketQueue


def synthetic_pocket_sizes(n_pockets: int, seed: int = 2024):
    """N_p ~ clip(N(330, 80), 150, 700)."""
    return np.clip(np.random.default_rng(seed).normal(330, 80, size=n_pockets), 150, 700).astype(int)


def _one_pocket(smp, perception, info, pid, n, B, timesteps, out_dir, score, dev):
    px, pt = synthetic.synthetic_pocket(1000 + pid, int(n))
    sizes = synthetic.synthetic_ligand_sizes(1000 + pid, B)
    onehot = np.eye(10, dtype=np.float32)[pt]
    pocket = {'x': torch.from_numpy(px).to(dev).repeat(B, 1), 'one_hot': torch.from_numpy(onehot).to(dev).repeat(B, 1),
              'size': torch.tensor([len(px)] * B, device=dev), 'mask': torch.arange(B, device=dev).repeat_interleave(len(px))}
    if score:
        pose = synthetic.synthetic_ligand_pose(1000 + pid, sizes, px.mean(axis=0, dtype=np.float64))
        pose[:, :3] -= px[0]
        smp.eps_transform = synthetic.PointMassScore(pose, smp.gamma, len(px), smp.T, dev)
    xh_lig, _, lig_mask, _ = smp.sample_given_pocket(pocket, sizes, timesteps=timesteps)
    mols = output.build_molecules(xh_lig[:, :3].contiguous(), xh_lig[:, 3:].argmax(1), lig_mask, B, info, perception)
    output.write_sdf_file(os.path.join(out_dir, f'warmup_{pid}.sdf'), [output.process_molecule(m, largest_frag=True) for m in mols])
    torch.cuda.synchronize()


def run_pocket_job(smp, perception, info, n_atoms, B: int, timesteps: int, out_dir: str, score: bool = True,
                   key: str = 'dndm/pocket_queue', warmup: bool = True):
    """Returns (seconds by wall clock -- max over ranks, barrier on both sides --, per-rank list of pocket records, seconds
    this rank spent waiting at the final barrier).  ``score``: install ``synthetic.PointMassScore`` per pocket so that the
    random-init denoiser keeps the ligands in the pocket (see its docstring); the pose is part of the synthetic pocket."""
    dev = smp.device
    world = dist.get_world_size() if dist.is_initialized() else 1
    queue = PocketQueue(len(n_atoms), costs=np.asarray(n_atoms, float), key=key)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    if warmup:          # one small untimed pocket: module loading, allocator growth, the first graph capture
        _one_pocket(smp, perception, info, -1, 150, min(B, 8), min(timesteps, 20), out_dir, score, dev)
    barrier()
    t0 = time.perf_counter()
    mine = []
    for pid in queue:
        t1 = time.perf_counter()
        px, pt = synthetic.synthetic_pocket(1000 + pid, int(n_atoms[pid]))
        sizes = synthetic.synthetic_ligand_sizes(1000 + pid, B)
        onehot = np.eye(10, dtype=np.float32)[pt]
        n = len(px)
        base_x, base_h = torch.from_numpy(px).to(dev), torch.from_numpy(onehot).to(dev)
        pocket = {'x': base_x.repeat(B, 1), 'one_hot': base_h.repeat(B, 1), 'size': torch.tensor([n] * B, device=dev),
                  'mask': torch.arange(B, device=dev).repeat_interleave(n)}
        if score:
            pose = synthetic.synthetic_ligand_pose(1000 + pid, sizes, px.mean(axis=0, dtype=np.float64))
            pose[:, :3] -= px[0]
            smp.eps_transform = synthetic.PointMassScore(pose, smp.gamma, n, smp.T, dev)
        torch.manual_seed(pid)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        xh_lig, xh_pocket, lig_mask, pocket_mask = smp.sample_given_pocket(pocket, sizes, timesteps=timesteps)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        mols = output.build_molecules(xh_lig[:, :3].contiguous(), xh_lig[:, 3:].argmax(1), lig_mask, B, info, perception)
        mols = [output.process_molecule(m, largest_frag=True) for m in mols]
        output.write_sdf_file(os.path.join(out_dir, f'pocket_{pid:04d}.sdf'), mols)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        e, el, ea = smp.engine.graph_stats_full()
        mine.append(dict(id=int(pid), atoms=int(n), s=round(t4 - t1, 3), setup_s=round(t2 - t1, 3), sample_s=round(t3 - t2, 3),
                         output_s=round(t4 - t3, 3), edges=int(e), ligand_receiver_share=round(el / max(e, 1), 3)))
    smp.eps_transform = None
    torch.cuda.synchronize()
    t_done = time.perf_counter()
    barrier()
    t_end = time.perf_counter()
    dt = torch.tensor([t_end - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, round(t_end - t_done, 3)))
    else:
        gathered = [(mine, 0.0)]
    return float(dt), [g[0] for g in gathered], [g[1] for g in gathered]
