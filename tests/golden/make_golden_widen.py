"""Generate ``widen.npz`` FROM THE REFERENCE ITSELF (build container only): forward cases of the OTHER pocket-conditional
configurations of the reference (widening), which the engine runs on its 256-channel kernels by exact
rewrites of the weight table (``diffndm_b200.weights.engine_table``):

* ``moad192``  -- the shape of configs/moad_fullatom_cond.yml: hidden_nf 192, edge_embedding_dim 8 (dynamics.py:118-127),
  pocket cutoff 4 A, interaction cutoff 7 A, joint_nf 128, six blocks;
* ``narrow128`` -- hidden_nf 128, joint_nf 32, five blocks (the network size of moad_ca_cond / crossdock_fullatom_joint), cutoffs 5 / 5;
* ``ca20``     -- residue_nf 20 != atom_nf 10 (C-alpha pockets: one node per residue, amino-acid one-hot; crossdock_ca_cond).

    python tests/golden/make_golden_widen.py

Each case stores the inputs, the reference's edge list, its fp64 output and the last block's h; weights are regenerated from
``random_init(cfg, seed)`` and guarded by a checksum."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_loader import build_reference_model  # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init, weights_checksum  # noqa: E402
from diffndm_b200 import synthetic  # noqa: E402

T = torch.from_numpy
npy = lambda t: t.detach().cpu().numpy().copy()

CASES = {
    'moad192': dict(cfg=dict(hidden_nf=192, edge_embedding_dim=8, edge_cutoff_pocket=4.0, edge_cutoff_interaction=7.0),
                    pocket=(31, 140), sizes=[9, 14, 6], seed=41, t=[0.7, 0.35, 0.05], spread=2.0),
    'narrow128': dict(cfg=dict(hidden_nf=128, joint_nf=32, n_layers=5), pocket=(32, 100), sizes=[11, 5], seed=42, t=[0.6, 0.2],
                      spread=2.0),
    'ca20': dict(cfg=dict(residue_nf=20), pocket=(33, 48), sizes=[12, 8, 15], seed=43, t=[0.8, 0.5, 0.1], spread=1.5),
}


def case_config(name):
    return DynamicsConfig(**CASES[name]['cfg'])


def main():
    torch.set_num_threads(8)
    out = {}
    for name, c in CASES.items():
        cfg = case_config(name)
        W = random_init(cfg, 7, 0.3)
        dyn64, _ = build_reference_model(cfg, W, dtype=torch.float64)
        px, pt = synthetic.synthetic_pocket(*c['pocket'])
        if cfg.residue_nf != 10:                                  # C-alpha-like pocket: sparser points, 20 residue types
            rng = np.random.default_rng(c['pocket'][0])
            px = (px * 1.8).astype(np.float32)
            pt = rng.integers(0, cfg.residue_nf, size=len(px))
        b = synthetic.make_batch(px, np.minimum(pt, 9), np.asarray(c['sizes']), c['seed'])
        if cfg.residue_nf != 10:                                  # rebuild the pocket features at the wider one-hot
            B, n_p = len(c['sizes']), len(px)
            oh = np.zeros((n_p, cfg.residue_nf), np.float32)
            oh[np.arange(n_p), pt] = 1.0
            b['xh_pocket'] = np.concatenate([b['xh_pocket'][:, :3], np.tile(oh, (B, 1))], 1).astype(np.float32)
        for s in range(len(c['sizes'])):
            m = b['lig_mask'] == s
            ctr = b['xh_lig'][m, :3].mean(0)
            b['xh_lig'][m, :3] = (b['xh_lig'][m, :3] - ctr) * c['spread'] + ctr
        t = np.asarray(c['t'], np.float32).reshape(-1, 1)
        trace = {}
        blk = dyn64.egnn._modules[f'e_block_{cfg.n_layers - 1}']
        hk = blk.register_forward_hook(lambda m, i, o: trace.__setitem__('h', npy(o[0])))
        with torch.no_grad():
            o64 = dyn64(T(b['xh_lig']).double(), T(b['xh_pocket']).double(), T(t).double(), T(b['lig_mask']), T(b['pocket_mask']))
            edges = dyn64.get_edges(T(b['lig_mask']), T(b['pocket_mask']), T(b['xh_lig'][:, :3]), T(b['xh_pocket'][:, :3]))
        hk.remove()
        n_l = len(b['lig_mask'])
        e = npy(edges)
        types = np.where((e[0] < n_l) & (e[1] < n_l), 1, np.where((e[0] >= n_l) & (e[1] >= n_l), 2, 0))
        o = dict(xh_lig=b['xh_lig'], xh_pocket=b['xh_pocket'], lig_mask=b['lig_mask'], pocket_mask=b['pocket_mask'], t=t,
                 out_lig_f64=npy(o64[0]), out_pocket_f64=npy(o64[1]), h_lig_last=trace['h'][:n_l].astype(np.float32),
                 edges=e.astype(np.int32), weights_checksum=np.asarray(weights_checksum(W)))
        print(f'{name}: N_l={n_l} N_p={len(b["pocket_mask"])} E={e.shape[1]} types={np.bincount(types, minlength=3).tolist()} '
              f'|eps_x|max={np.abs(o["out_lig_f64"][:, :3]).max():.3f} |eps_h|max={np.abs(o["out_lig_f64"][:, 3:]).max():.3f}')
        for k, v in o.items():
            out[f'{name}/{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'widen.npz'), **out)
    print('wrote widen.npz', os.path.getsize(os.path.join(HERE, 'widen.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
