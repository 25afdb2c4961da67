"""BASELINE configs[1] as a whole job: P synthetic CrossDocked-shaped pockets (sizes ~ clip(N(330, 80), 150, 700)) x B ligands,
500 steps, pockets pulled by the ranks from the shared ``PocketQueue`` (largest first), one SDF file per pocket -- the
layout of the reference's ``my_test.py`` (one process per pocket, one ``<pocket>.sdf`` each).

    python scripts/sample_pockets.py [--pockets 8] [--batch 100] [--timesteps 500] [--out DIR]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/sample_pockets.py ...

Rank 0 prints one JSON object: whole-job ligands/s by wall clock (max over ranks, barrier on both sides; includes graph
capture per new batch shape, bond perception, SDF writing), pockets per rank, and the per-pocket seconds."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, output, synthetic                   # noqa: E402
from diffndm_b200.chem import BondPerception                              # noqa: E402
from diffndm_b200.datasets import crossdock_dataset_info                  # noqa: E402
from diffndm_b200.job import run_pocket_job, synthetic_pocket_sizes       # noqa: E402
from diffndm_b200.sampler import ConditionalSampler                       # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pockets', type=int, default=8)
    ap.add_argument('--batch', type=int, default=100)
    ap.add_argument('--timesteps', type=int, default=500)
    ap.add_argument('--out', type=str, default=None)
    ap.add_argument('--no-score', action='store_true', help='plain random-init denoiser (ligands leave the pocket: cheaper, unrepresentative steps)')
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    out_dir = args.out or tempfile.mkdtemp()
    os.makedirs(out_dir, exist_ok=True)
    B = args.batch
    n_atoms = synthetic_pocket_sizes(args.pockets)
    cfg = DynamicsConfig()
    info = crossdock_dataset_info()
    nmax = int(n_atoms.max())
    dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=B * (nmax + 60) + 1024,
                             max_edges=B * (nmax + 60) * 48, max_samples=max(B, 8)).eval()
    smp = ConditionalSampler(dyn, timesteps=500)
    perception = BondPerception(dyn.engine, info)
    dt, gathered, idle = run_pocket_job(smp, perception, info, n_atoms, B, args.timesteps, out_dir, score=not args.no_score)
    if rank == 0:
        done = sorted((p for g in gathered for p in g), key=lambda p: p['id'])
        assert [p['id'] for p in done] == list(range(args.pockets)), 'every pocket exactly once'
        print(json.dumps({'pockets': args.pockets, 'ligands_per_pocket': B, 'timesteps': args.timesteps, 'n_gpus': world,
                          'seconds': round(float(dt), 3),
                          'ligands_per_s': round(args.pockets * B * (500 / args.timesteps) / float(dt), 2),
                          'pockets_per_rank': [len(g) for g in gathered], 'tail_idle_s_per_rank': idle,
                          'per_pocket': done, 'flags': dyn.engine.read_flags()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
