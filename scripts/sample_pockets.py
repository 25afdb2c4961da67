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
from diffndm_b200.parallel import PocketQueue                             # noqa: E402
from diffndm_b200.sampler import ConditionalSampler                       # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pockets', type=int, default=8)
    ap.add_argument('--batch', type=int, default=100)
    ap.add_argument('--timesteps', type=int, default=500)
    ap.add_argument('--out', type=str, default=None)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    out_dir = args.out or tempfile.mkdtemp()
    os.makedirs(out_dir, exist_ok=True)
    B = args.batch
    rng = np.random.default_rng(2024)
    n_atoms = np.clip(rng.normal(330, 80, size=args.pockets), 150, 700).astype(int)
    cfg = DynamicsConfig()
    info = crossdock_dataset_info()
    nmax = int(n_atoms.max())
    dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 1e-3), max_nodes=B * (nmax + 60) + 1024,
                             max_edges=B * (nmax + 60) * 44, max_samples=max(B, 8)).eval()
    smp = ConditionalSampler(dyn, timesteps=500)
    perception = BondPerception(dyn.engine, info)
    queue = PocketQueue(args.pockets, costs=n_atoms.astype(float) ** 1.0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

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
        torch.manual_seed(pid)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        xh_lig, xh_pocket, lig_mask, pocket_mask = smp.sample_given_pocket(pocket, sizes, timesteps=args.timesteps)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        mols = output.build_molecules(xh_lig[:, :3].contiguous(), xh_lig[:, 3:].argmax(1), lig_mask, B, info, perception)
        mols = [output.process_molecule(m, largest_frag=True) for m in mols]
        output.write_sdf_file(os.path.join(out_dir, f'pocket_{pid:04d}.sdf'), mols)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        mine.append((int(pid), int(n), round(t4 - t1, 3), round(t2 - t1, 3), round(t3 - t2, 3), round(t4 - t3, 3),
                     int(dyn.engine.graph_stats()[0])))
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    else:
        gathered = [mine]
    if rank == 0:
        done = sorted(p for g in gathered for p in g)
        assert [p[0] for p in done] == list(range(args.pockets)), 'every pocket exactly once'
        print(json.dumps({'pockets': args.pockets, 'ligands_per_pocket': B, 'timesteps': args.timesteps, 'n_gpus': world,
                          'seconds': round(float(dt), 3),
                          'ligands_per_s': round(args.pockets * B * (500 / args.timesteps) / float(dt), 2),
                          'pockets_per_rank': [len(g) for g in gathered],
                          'per_pocket': [{'id': p[0], 'atoms': p[1], 's': p[2], 'setup_s': p[3], 'sample_s': p[4], 'output_s': p[5], 'edges': p[6]}
                                         for p in done],
                          'flags': dyn.engine.read_flags()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
