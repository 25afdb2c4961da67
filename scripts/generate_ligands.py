"""``generate_ligands.py`` of the reference (generate_ligands.py:20-108) on the B200 engine: same positional argument and
flags, same batch loop, one SDF file out.

    python scripts/generate_ligands.py <checkpoint.ckpt> --pdbfile P --ref_ligand A:330 --outfile out.sdf --n_samples 20

Differences, all forced by what the path is: the checkpoint is read for its ``ddpm.dynamics.*`` weights and the ligand-size
histogram only (no Lightning module); ``--sanitize`` / ``--relax`` need RDKit and are rejected unless a host ``mol_builder``
is plugged in; ``--optimize/--path/--path_save`` (the RL noise-adjust net) are out of scope;
``--SVDD 1`` / ``--SPSA 1`` need a host reward (``--reward module:function`` giving
``reward_fn(x_lig, atom_types, lig_mask) -> list[float]``).  ``--random_init`` replaces the checkpoint by seeded random
weights (smoke tests: there is no network to fetch the Zenodo checkpoint from)."""
import argparse
import importlib
import os
import sys
from pathlib import Path

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200.datasets import crossdock_dataset_info                     # noqa: E402
from diffndm_b200.engine import B200EGNNDynamics                             # noqa: E402
from diffndm_b200.generate import LigandGenerator  # noqa: E402
from diffndm_b200.sampler import ConditionalSampler                          # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init                 # noqa: E402


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('checkpoint', type=Path, nargs='?', default=None)
    parser.add_argument('--pdbfile', type=str, required=True)
    parser.add_argument('--resi_list', type=str, nargs='+', default=None)
    parser.add_argument('--ref_ligand', type=str, default=None)
    parser.add_argument('--outfile', type=Path, required=True)
    parser.add_argument('--n_samples', type=int, default=20)
    parser.add_argument('--batch_size', type=int, default=None)
    parser.add_argument('--num_nodes_lig', type=int, default=None)
    parser.add_argument('--all_frags', action='store_true')
    parser.add_argument('--sanitize', action='store_true')
    parser.add_argument('--relax', action='store_true')
    parser.add_argument('--resamplings', type=int, default=10)       # accepted for compatibility (inpainting only)
    parser.add_argument('--jump_length', type=int, default=1)        # accepted for compatibility (inpainting only)
    parser.add_argument('--timesteps', type=int, default=None)
    parser.add_argument('--SVDD', type=int, default=0)
    parser.add_argument('--SPSA', type=int, default=0)
    parser.add_argument('--reward', type=str, default=None, help='module:function of the host reward for --SVDD / --SPSA')
    parser.add_argument('--random_init', type=int, default=None, help='seed of random weights instead of a checkpoint')
    parser.add_argument('--seed', type=int, default=None)
    args = parser.parse_args(argv)
    if (args.checkpoint is None) == (args.random_init is None):
        parser.error('give a checkpoint or --random_init SEED')
    if args.batch_size is None:
        args.batch_size = args.n_samples
    assert args.n_samples % args.batch_size == 0                      # generate_ligands.py:52
    if args.seed is not None:
        torch.manual_seed(args.seed)
        torch.cuda.manual_seed(args.seed)

    reward_fn = None
    if args.reward:
        mod, fn = args.reward.split(':')
        reward_fn = getattr(importlib.import_module(mod), fn)
    if (args.SVDD or args.SPSA) and reward_fn is None:
        parser.error('--SVDD / --SPSA need --reward module:function (host chemistry stays external)')
    if args.checkpoint is not None:      # sizes, cutoffs, schedule and the size histogram come out of the checkpoint
        model = LigandGenerator.from_checkpoint(args.checkpoint)
    else:
        cfg = DynamicsConfig()
        dyn = B200EGNNDynamics(cfg, random_init(cfg, args.random_init, 1e-3)).eval()
        model = LigandGenerator(ConditionalSampler(dyn, timesteps=500), crossdock_dataset_info())
    n = model.generate_to_sdf(args.pdbfile, args.outfile, n_samples=args.n_samples, batch_size=args.batch_size,
                              num_nodes_lig=args.num_nodes_lig, all_frags=args.all_frags, pocket_ids=args.resi_list,
                              ref_ligand=args.ref_ligand, sanitize=args.sanitize, relax_iter=(200 if args.relax else 0),
                              timesteps=args.timesteps, svdd=args.SVDD, spsa=args.SPSA, reward_fn=reward_fn)
    print(f'wrote {n} molecules to {args.outfile}')
    return n


if __name__ == '__main__':
    main()
