"""Exact dead-work elimination (csrc/graph.cuh, engine.cu) as a property of the reference algorithm, checked on the oracle in
fp64: when only the ligand output is read, block n_layers - 1 - k only has to aggregate messages for the receivers S_k,
S_0 = ligand atoms + their senders, S_k = S_{k-1} + all senders of S_{k-1} (the graph is the same in every block,
dynamics.py:114).  The engine applies levels k = 0, 1 (DNDM_PRUNE_LEVELS); the GPU test test_last_block_pruning_is_exact checks
its bits, this test checks the argument itself -- for every number of levels, and that the sets are not looser than needed."""
import numpy as np
import pytest

from diffndm_b200 import synthetic
from diffndm_b200.weights import DynamicsConfig, random_init
from oracle import egnn_oracle as O


def _forward(W, b, t, cfg, levels, shrink=0):
    """O.dynamics_forward with the trailing ``levels`` GCLs restricted to the receivers of their hop set (``shrink`` = 1 uses
    the set of the NEXT block instead: one hop too few)."""
    dt = np.float64
    Wd = {k: v.astype(dt) for k, v in W.items()}
    n_l = len(b['lig_mask'])
    mask = np.concatenate([b['lig_mask'], b['pocket_mask']])
    x = np.concatenate([b['xh_lig'][:, :3], b['xh_pocket'][:, :3]]).astype(dt)
    h_l = O.linear(O.silu(O.linear(b['xh_lig'][:, 3:].astype(dt), Wd['atom_encoder.0.weight'], Wd['atom_encoder.0.bias'])),
                   Wd['atom_encoder.2.weight'], Wd['atom_encoder.2.bias'])
    h_p = O.linear(O.silu(O.linear(b['xh_pocket'][:, 3:].astype(dt), Wd['residue_encoder.0.weight'], Wd['residue_encoder.0.bias'])),
                   Wd['residue_encoder.2.weight'], Wd['residue_encoder.2.bias'])
    h = np.concatenate([np.concatenate([h_l, h_p]), np.asarray(t, dt).reshape(-1, 1)[mask]], axis=1)
    edges = O.get_edges(b['lig_mask'], b['pocket_mask'], b['xh_lig'][:, :3], b['xh_pocket'][:, :3], cfg)
    row, col = edges[0], edges[1]
    N = len(mask)
    # hop sets: S[0] = ligand atoms and their senders, S[k] = S[k-1] and all its senders
    S = []
    cur = np.zeros(N, bool)
    cur[:n_l] = True
    for _ in range(levels + 1):
        nxt = cur.copy()
        nxt[col[cur[row]]] = True
        S.append(nxt)
        cur = nxt
    upd = np.concatenate([np.ones(n_l, dt), np.zeros(N - n_l, dt)])[:, None]
    r0, _ = O.coord2diff(x, row, col, 1)
    h = O.linear(h, Wd['egnn.embedding.weight'], Wd['egnn.embedding.bias'])
    x_cur = x
    shares = []
    for i in range(cfg.n_layers):
        p = f'egnn.e_block_{i}.'
        radial, cdiff = O.coord2diff(x_cur, row, col, cfg.norm_constant)
        ccross = O.coord2cross(x_cur, row, col, mask, cfg.norm_constant)
        attr = np.concatenate([radial, r0], axis=1)
        k = cfg.n_layers - 1 - i
        if k < levels:
            if not shrink:
                keep = S[k][row]
            else:                                           # one hop too few: the set of the block after this one
                keep = S[k - 1][row] if k >= 1 else (row < n_l)
            shares.append(keep.mean())
            h, _ = O.gcl_forward(Wd, p + 'gcl_0.', h, row[keep], col[keep], attr[keep], cfg)
        else:
            h, _ = O.gcl_forward(Wd, p + 'gcl_0.', h, row, col, attr, cfg)
        lig = row < n_l                                     # the coordinate heads only move ligand atoms
        x_cur, _, _ = O.equiv_forward(Wd, p + 'gcl_equiv.', h, x_cur, row[lig], col[lig], cdiff[lig], ccross[lig], attr[lig], upd, cfg)
    h = O.linear(h, Wd['egnn.embedding_out.weight'], Wd['egnn.embedding_out.bias'])[:, :-1]
    h_fa = O.linear(O.silu(O.linear(h[:n_l], Wd['atom_decoder.0.weight'], Wd['atom_decoder.0.bias'])),
                    Wd['atom_decoder.2.weight'], Wd['atom_decoder.2.bias'])
    return np.concatenate([(x_cur - x)[:n_l], h_fa], axis=1), shares


@pytest.fixture(scope='module')
def case():
    cfg = DynamicsConfig()
    W = random_init(cfg, 3, 0.3)
    px, pt = synthetic.synthetic_pocket(77, 160)
    b = synthetic.make_batch(px, pt, np.array([9, 13]), 4)
    t = np.array([[0.6], [0.3]], np.float32)
    ocfg = O.OracleConfig()
    ref, _ = O.dynamics_forward(W, b['xh_lig'], b['xh_pocket'], t, b['lig_mask'], b['pocket_mask'], ocfg, dtype=np.float64)
    return W, b, t, ocfg, ref


@pytest.mark.parametrize('levels', [0, 1, 2, 3, 4])
def test_pruned_trailing_blocks_leave_the_ligand_output_unchanged(case, levels):
    W, b, t, cfg, ref = case
    out, shares = _forward(W, b, t, cfg, levels)
    assert np.abs(out - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    assert all(s1 >= s0 for s0, s1 in zip(shares[::-1][:-1], shares[::-1][1:]))       # the sets grow block by block backwards
    if levels:
        assert shares[-1] < 0.75                                                        # the last block really drops edges


def test_one_hop_too_few_changes_the_output(case):
    W, b, t, cfg, ref = case
    out, _ = _forward(W, b, t, cfg, 2, shrink=1)
    assert np.abs(out - ref).max() > 1e-6
