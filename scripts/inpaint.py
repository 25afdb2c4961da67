"""``inpaint.py`` of the reference (inpaint.py:192-236) on the B200 engine: same positional argument and flags, one SDF out.

    python scripts/inpaint.py <checkpoint.ckpt> --pdbfile P --ref_ligand A:330 --fix_atoms C1 N6 C5 --outfile out.sdf

``--fix_atoms`` takes atom names of the PDB ligand residue or SDF files (inpaint.py:47-62).  ``--save_traj`` (visualisation) is
not offered; ``--sanitize`` / ``--relax`` need RDKit (rejected without a host ``mol_builder``); ``--svdd 1`` needs
``--reward module:function``.  ``--random_init SEED`` replaces the checkpoint (smoke tests)."""
import argparse
import importlib
import os
import sys
from pathlib import Path

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import output                                              # noqa: E402
from diffndm_b200.datasets import crossdock_dataset_info                     # noqa: E402
from diffndm_b200.engine import B200EGNNDynamics                             # noqa: E402
from diffndm_b200.generate import LigandGenerator  # noqa: E402
from diffndm_b200.sampler import ConditionalSampler                          # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init                 # noqa: E402


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('checkpoint', type=Path, nargs='?', default=None)
    parser.add_argument('--pdbfile', type=str, required=True)
    parser.add_argument('--ref_ligand', type=str, required=True)
    parser.add_argument('--fix_atoms', type=str, nargs='+', required=True)
    parser.add_argument('--center', type=str, default='ligand', choices={'ligand', 'pocket'})
    parser.add_argument('--outfile', type=Path, required=True)
    parser.add_argument('--n_samples', type=int, default=20)
    parser.add_argument('--add_n_nodes', type=int, default=None)
    parser.add_argument('--relax', action='store_true')
    parser.add_argument('--sanitize', action='store_true')
    parser.add_argument('--resamplings', type=int, default=20)
    parser.add_argument('--timesteps', type=int, default=50)
    parser.add_argument('--svdd', type=int, default=0)
    parser.add_argument('--reward', type=str, default=None, help='module:function of the host reward for --svdd')
    parser.add_argument('--random_init', type=int, default=None, help='seed of random weights instead of a checkpoint')
    parser.add_argument('--seed', type=int, default=None)
    args = parser.parse_args(argv)
    if (args.checkpoint is None) == (args.random_init is None):
        parser.error('give a checkpoint or --random_init SEED')
    if args.seed is not None:
        torch.manual_seed(args.seed)
        torch.cuda.manual_seed(args.seed)
    reward_fn = None
    if args.reward:
        mod, fn = args.reward.split(':')
        reward_fn = getattr(importlib.import_module(mod), fn)
    if args.svdd and reward_fn is None:
        parser.error('--svdd needs --reward module:function (host chemistry stays external)')
    if args.checkpoint is not None:      # sizes, cutoffs, schedule and the size histogram come out of the checkpoint
        model = LigandGenerator.from_checkpoint(args.checkpoint)
    else:
        cfg = DynamicsConfig()
        dyn = B200EGNNDynamics(cfg, random_init(cfg, args.random_init, 1e-3)).eval()
        model = LigandGenerator(ConditionalSampler(dyn, timesteps=500), crossdock_dataset_info())
    molecules = model.inpaint_ligand(args.pdbfile, args.n_samples, args.ref_ligand, args.fix_atoms, args.add_n_nodes,
                                     args.svdd, center=args.center, sanitize=args.sanitize, largest_frag=False,
                                     relax_iter=(200 if args.relax else 0), timesteps=args.timesteps,
                                     resamplings=args.resamplings, reward_fn=reward_fn)
    n = output.write_sdf_file(args.outfile, molecules)
    print(f'wrote {n} molecules to {args.outfile}')
    return n


if __name__ == '__main__':
    main()
