"""Run-to-run determinism on one GPU: the denoiser forward, a graphed trajectory and an eager trajectory, each twice from
the same seed.  Prints one JSON object (bitwise equality per case, max abs difference otherwise)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, synthetic                          # noqa: E402
from diffndm_b200.sampler import ConditionalSampler                      # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init             # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    cfg = DynamicsConfig()
    B = 20
    px, pt = synthetic.synthetic_pocket(3, 200)
    sizes = synthetic.synthetic_ligand_sizes(3, B)
    b = synthetic.make_batch(px, pt, sizes, 3)
    N = len(b['lig_mask']) + len(b['pocket_mask'])
    dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=N + 256, max_edges=N * 48, max_samples=B).eval()
    t = lambda a: torch.from_numpy(a).to(dev)
    args = (t(b['xh_lig']), t(b['xh_pocket']), torch.full((B, 1), 0.4, device=dev), t(b['lig_mask']), t(b['pocket_mask']))
    outs = [dyn(*args)[0].clone() for _ in range(20)]
    res = {'forward_20x_bitwise': all(torch.equal(outs[0], o) for o in outs)}
    smp = ConditionalSampler(dyn, timesteps=500)
    oh = np.eye(10, dtype=np.float32)[pt]
    mk = lambda: {'x': t(np.tile(px, (B, 1))), 'one_hot': t(np.tile(oh, (B, 1))), 'size': torch.tensor([len(px)] * B),
                  'mask': torch.arange(B).repeat_interleave(len(px))}
    pose = synthetic.synthetic_ligand_pose(3, sizes, px.mean(0))
    pose[:, :3] -= px[0]
    smp.eps_transform = synthetic.PointMassScore(pose, smp.gamma, len(px), 500, dev)
    for name, graph in (('eager', False), ('graph', True)):
        runs = []
        for rep in range(3):
            torch.manual_seed(11)
            runs.append(smp.sample_given_pocket(mk(), sizes, timesteps=60, use_cuda_graph=graph)[0].clone())
        res[f'{name}_trajectory_bitwise'] = [torch.equal(runs[0], r) for r in runs[1:]]
        res[f'{name}_trajectory_max_diff'] = [float((runs[0] - r).abs().max()) for r in runs[1:]]
    print(json.dumps(res))


if __name__ == '__main__':
    main()
