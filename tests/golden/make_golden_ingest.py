"""Generate ``ingest.npz`` FROM THE REFERENCE ITSELF (build container only).

    python tests/golden/make_golden_ingest.py

Runs the unmodified ``DistributionNodes`` (equivariant_diffusion/en_diffusion.py:963-1033), ``utils.num_nodes_to_batch_mask``
and ``utils.batch_to_list`` (utils.py:130-153) on seeded inputs and stores their outputs.  BioPython is not installed, so the
PDB-reading half of the ingest (``get_pocket_from_ligand`` / ``prepare_pocket``) has no reference run; it is pinned by the
atom count SURVEY.md section 8d quotes for ``example/3rfm.pdb`` (286) and by hand-written PDB records in the tests.
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

import ref_loader  # noqa: E402


def main():
    ref_loader.install_shims()
    with contextlib.redirect_stdout(io.StringIO()):
        from equivariant_diffusion.en_diffusion import DistributionNodes
        import utils as ref_utils
    rng = np.random.default_rng(7)
    # CrossDocked-like joint histogram: ligand sizes 0..49, pocket sizes 0..63 (sparse, with empty rows and columns)
    hist = np.zeros((50, 64), np.float64)
    for _ in range(4000):
        n2 = int(np.clip(rng.normal(40, 9), 8, 63))
        n1 = int(np.clip(rng.normal(0.45 * n2 + 5, 5), 4, 49))
        hist[n1, n2] += 1
    out = {'histogram': hist}
    with contextlib.redirect_stdout(io.StringIO()):
        dist = DistributionNodes(hist)
    n2 = torch.tensor(rng.integers(8, 64, size=40))
    n1 = torch.tensor(rng.integers(4, 50, size=40))
    torch.manual_seed(1234)
    out['cond_n2'] = n2.numpy()
    out['sample_n1_given_n2'] = dist.sample_conditional(n1=None, n2=n2).numpy()
    out['cond_n1'] = n1.numpy()
    out['sample_n2_given_n1'] = dist.sample_conditional(n1=n1, n2=None).numpy()
    a, b = dist.sample(25)
    out['joint_n1'], out['joint_n2'] = a.numpy(), b.numpy()
    out['log_prob'] = dist.log_prob(n1, n2).numpy()
    out['log_prob_n1_given_n2'] = dist.log_prob_n1_given_n2(n1, n2).numpy()
    out['log_prob_n2_given_n1'] = dist.log_prob_n2_given_n1(n2, n1).numpy()
    out['entropy'] = np.float64(dist.m.entropy().item())

    sizes = torch.tensor([3, 1, 4, 1, 5])
    mask = ref_utils.num_nodes_to_batch_mask(5, sizes, 'cpu')
    out['mask_sizes'] = sizes.numpy()
    out['mask'] = mask.numpy()
    out['mask_int'] = ref_utils.num_nodes_to_batch_mask(4, 3, 'cpu').numpy()
    perm = torch.tensor(rng.permutation(len(mask)))
    data = torch.arange(len(mask) * 2, dtype=torch.float32).view(-1, 2)
    chunks = ref_utils.batch_to_list(data[perm], mask[perm])
    out['b2l_perm'] = perm.numpy()
    out['b2l_sizes'] = np.array([len(c) for c in chunks])
    out['b2l_sorted_rows'] = np.concatenate([np.sort(c[:, 0].numpy()) for c in chunks])
    np.savez_compressed(os.path.join(HERE, 'ingest.npz'), **out)
    print('wrote ingest.npz', {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == '__main__':
    main()
