"""Generate ``bonds.npz`` FROM THE REFERENCE ITSELF (build container only).

    python tests/golden/make_golden_bonds.py

Runs the unmodified ``analysis.molecule_builder.get_bond_order_batch`` exactly as ``make_mol_edm`` calls it
(molecule_builder.py:100-113: torch.cdist distances, cartesian product of the atom types, the crossdock distance tables of
``constants.dataset_params``) on seeded molecule-like point sets and stores the resulting lower-triangular bond-order
matrices together with the tables.  Also records, per case, how many atom pairs lie within 1e-3 pm of a decision
threshold (where cdist's and the direct-difference arithmetic could legitimately disagree).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402


def molecule(rng, n, n_types):
    """Chain-and-branch growth with chemically plausible steps (1.2-1.6 A) plus a few stray atoms."""
    pts = [np.zeros(3)]
    while len(pts) < n:
        base = pts[rng.integers(len(pts))]
        step = rng.normal(size=3)
        step *= rng.uniform(1.15, 1.6) / np.linalg.norm(step)
        cand = base + step
        if min(np.linalg.norm(cand - q) for q in pts) > 1.05:
            pts.append(cand)
    x = np.asarray(pts)
    if n > 6:
        x[-1] += rng.normal(size=3) * 3.0               # a detached atom -> more than one fragment
    t = rng.choice(n_types, size=n, p=None if n_types != 10 else [.62, .12, .17, .02, 0, .01, .02, .02, 0, .02])
    return np.round(x, 3).astype(np.float32), t.astype(np.int64)


def main():
    ref_loader.install_shims()
    import constants                                        # noqa: E402  (reference)
    from analysis.molecule_builder import get_bond_order_batch  # noqa: E402  (reference)
    info = constants.dataset_params['crossdock']
    b1, b2, b3 = (np.asarray(info[k], np.float32) for k in ('bonds1', 'bonds2', 'bonds3'))
    margins = np.asarray([constants.margin1, constants.margin2, constants.margin3], np.float32)
    rng = np.random.default_rng(11)
    out = dict(bonds1=b1, bonds2=b2, bonds3=b3, margins=margins)
    for name, sizes in [('mix_b6', [5, 23, 31, 12, 50, 8]), ('small_b3', [1, 2, 9])]:
        xs, ts, mask, mats, near = [], [], [], [], 0
        for b, n in enumerate(sizes):
            x, t = molecule(rng, n, 10)
            pos = torch.from_numpy(x)
            dists = torch.cdist(pos.unsqueeze(0), pos.unsqueeze(0), p=2).squeeze(0).view(-1)     # :107-108
            a1, a2 = torch.cartesian_prod(torch.from_numpy(t), torch.from_numpy(t)).T             # :109
            e_full = get_bond_order_batch(a1, a2, dists, info).view(n, n)                         # :110
            e = torch.tril(e_full, diagonal=-1)                                                   # :111
            d = 100.0 * dists.view(n, n).double().numpy()
            for tab, m in ((b1, margins[0]), (b2, margins[1]), (b3, margins[2])):
                thr = tab[t[:, None], t[None, :]].astype(np.float64) + float(m)
                near += int((np.abs(d - thr)[np.tril_indices(n, -1)] < 1e-3).sum())
            xs.append(x); ts.append(t); mask.append(np.full(n, b, np.int64)); mats.append(e.numpy().astype(np.int8).reshape(-1))
        out[f'{name}/x'] = np.concatenate(xs)
        out[f'{name}/types'] = np.concatenate(ts)
        out[f'{name}/mask'] = np.concatenate(mask)
        out[f'{name}/sizes'] = np.asarray(sizes, np.int64)
        out[f'{name}/e_flat'] = np.concatenate(mats)          # per molecule n*n int8, concatenated
        out[f'{name}/pairs_near_threshold'] = np.int64(near)
        print(f'bonds[{name}]: atoms={sum(sizes)} bonds={int((np.concatenate(mats) > 0).sum())} near-threshold pairs={near}')
    np.savez_compressed(os.path.join(HERE, 'bonds.npz'), **out)
    print('bonds.npz', os.path.getsize(os.path.join(HERE, 'bonds.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
