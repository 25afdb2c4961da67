"""Generate ``distribution.npz`` FROM THE REFERENCE ITSELF (build container only): end-of-trajectory ligands of FULL
500-step free-running ``ConditionalDDPM.sample_given_pocket`` runs (its own Gaussian draws, not recorded).

    python tests/golden/make_golden_distribution.py [n_batches]

SURVEY.md section 8c-(iv): the trajectory bar is distributional ("end-of-trajectory ... distributions must be
statistically indistinguishable").  RDKit / OpenBabel are not installed, so QED / SA cannot be evaluated on either side;
the fixture stores the final ligands themselves and ``tests/test_gpu_parity.py::test_trajectory_distribution_vs_reference``
compares the RDKit-free statistics the survey names (atom-type histogram, pairwise-distance distribution, radius of
gyration, distance to the pocket, EDM bond counts) between these and ligands the CUDA engine draws with its own noise.
Takes a few minutes of CPU time.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_loader import build_reference_model  # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init, weights_checksum  # noqa: E402
from diffndm_b200 import synthetic  # noqa: E402

SIZES = [9, 12, 10, 14, 8, 11, 13, 10, 12, 9, 11, 10, 14, 8, 12, 13]      # one batch = 16 ligands, fixed sizes
POCKET_SEED, POCKET_ATOMS = 41, 48


def main():
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.set_num_threads(8)
    cfg = DynamicsConfig()
    seed_w, gain = 0, 0.3
    W = random_init(cfg, seed_w, gain)
    dyn, ddpm = build_reference_model(cfg, W)
    ddpm.handle_to_mol = lambda *a, **k: [[]]
    ddpm.my_reward_function = lambda *a, **k: 0.0
    px, pt = synthetic.synthetic_pocket(POCKET_SEED, POCKET_ATOMS)
    B, n_p = len(SIZES), len(px)
    onehot = np.eye(cfg.atom_nf, dtype=np.float32)[pt]
    finals = []
    for r in range(n_batches):
        torch.manual_seed(9000 + r)
        pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))),
                  'size': torch.tensor([n_p] * B), 'mask': torch.arange(B).repeat_interleave(n_p)}
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            xh_lig, xh_pocket, lig_mask, pocket_mask = ddpm.sample_given_pocket(
                pocket, torch.tensor(SIZES), torch.zeros(B, 3), None, False, 0, False, 'x', 'cpu',
                0, None, None, 0, 0, timesteps=500)
        # ligand coordinates relative to the (translated) pocket's centre of mass: translation-free
        per = []
        for b in range(B):
            sel = (lig_mask == b).numpy()
            pcb = xh_pocket[pocket_mask == b][:, :3].mean(0).numpy()
            per.append(xh_lig[:, :3].numpy()[sel] - pcb)
        finals.append((np.concatenate(per), xh_lig[:, 3:].argmax(1).numpy()))
        print(f'batch {r}: {time.time() - t0:.1f} s, |x|max={np.abs(finals[-1][0]).max():.2f}, '
              f'types={np.bincount(finals[-1][1], minlength=10).tolist()}', flush=True)
    out = dict(pocket_x=px, pocket_t=pt, sizes=np.asarray(SIZES), n_batches=np.int64(n_batches),
               x_rel=np.stack([f[0] for f in finals]).astype(np.float32), types=np.stack([f[1] for f in finals]),
               weight_seed=seed_w, coord_head_gain=gain, weights_checksum=weights_checksum(W))
    np.savez_compressed(os.path.join(HERE, 'distribution.npz'), **out)
    print('wrote distribution.npz', os.path.getsize(os.path.join(HERE, 'distribution.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
