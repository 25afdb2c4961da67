import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from diffndm_b200 import engine as E, synthetic
from diffndm_b200 import sampler as S
from diffndm_b200.weights import DynamicsConfig, random_init
B, n_p = 100, 330
cfg = DynamicsConfig()
dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=B * (n_p + 50) + 1024, max_edges=B * (n_p + 50) * 40, max_samples=B).eval()
px, pt = synthetic.synthetic_pocket(7, n_p)
sizes = synthetic.synthetic_ligand_sizes(7, B)
onehot = np.eye(10, dtype=np.float32)[pt]
dev = torch.device('cuda')
pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))).to(dev), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))).to(dev),
          'size': torch.tensor([len(px)] * B, device=dev), 'mask': torch.arange(B, device=dev).repeat_interleave(len(px))}
smp = S.ConditionalSampler(dyn, timesteps=500)
pose = synthetic.synthetic_ligand_pose(7, sizes, px.mean(axis=0, dtype=np.float64)); pose[:, :3] -= px[0]
smp.eps_transform = synthetic.PointMassScore(pose, smp.gamma, len(px), smp.T, dev)
def run():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    smp.sample_given_pocket(pocket, sizes, timesteps=500)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return round(t1 - t0, 4), round(t2 - t0, 4)
print('warm', run()); print('normal (return, synced)', run(), run())
orig = S._GraphedReverseStep.__call__
def fake(self, s):
    self._next_s = s - 1
    self.engine.set_static_masks(True)
S._GraphedReverseStep.__call__ = fake
print('no replay: host loop only', run(), run())
S._GraphedReverseStep.__call__ = orig
import cProfile, pstats
S._GraphedReverseStep.__call__ = fake
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
