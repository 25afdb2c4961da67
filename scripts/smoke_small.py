"""Small end-to-end pass (handy under a debugger or a memory checker): forward with and without pocket output, radius graph, sampler
step, a 3-step trajectory and a short inpainting run on a tiny batch."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, synthetic
from diffndm_b200.sampler import ConditionalSampler
from diffndm_b200.weights import DynamicsConfig, random_init

dev = torch.device('cuda')
cfg = DynamicsConfig()
dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=2048, max_edges=100000, max_samples=16).eval()
px, pt = synthetic.synthetic_pocket(3, 120)
sizes = np.array([7, 23, 5, 31])
b = synthetic.make_batch(px, pt, sizes, 3)
t = lambda a: torch.from_numpy(a).to(dev)
B = len(sizes)
args = (t(b['xh_lig']), t(b['xh_pocket']), torch.full((B, 1), 0.4, device=dev), t(b['lig_mask']), t(b['pocket_mask']))
ol, op = dyn(*args, n_samples=B)
dyn.compute_pocket_output = False
ol2, _ = dyn(*args, n_samples=B)
torch.cuda.synchronize()
print('forward ok', float(ol.abs().max()), float((ol - ol2).abs().max()))
rp, col = dyn.engine.radius_graph(args[0], args[1], args[3], args[4], B)
print('graph ok', int(rp[-1]), col.shape)
smp = ConditionalSampler(dyn, timesteps=500)
onehot = np.eye(10, dtype=np.float32)[pt]
pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))),
          'size': torch.tensor([len(px)] * B), 'mask': torch.arange(B).repeat_interleave(len(px))}
xh, xp, lm, pm = smp.sample_given_pocket(pocket, sizes, timesteps=3)
print('trajectory ok', xh.shape, float(xh[:, :3].abs().max()))
rng = np.random.default_rng(0)
lig = {'x': torch.from_numpy((px.mean(0)[None] + rng.normal(size=(int(sizes.sum()), 3))).astype(np.float32)),
       'one_hot': torch.from_numpy(np.eye(10, dtype=np.float32)[rng.integers(0, 10, int(sizes.sum()))]),
       'size': torch.from_numpy(sizes), 'mask': torch.from_numpy(np.repeat(np.arange(B), sizes))}
fixed = torch.from_numpy(np.concatenate([(np.arange(n) < 3) for n in sizes]).astype(np.float32))
xh, xp, lm, pm = smp.inpaint(lig, pocket, fixed, resamplings=2, timesteps=2)
torch.cuda.synchronize()
print('inpaint ok', xh.shape)
