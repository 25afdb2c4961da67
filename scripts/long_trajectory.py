"""500-step free-running sampling on one synthetic pocket: flags must stay clean, the state finite and COM-free."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, synthetic
from diffndm_b200.sampler import ConditionalSampler
from diffndm_b200.chem import BondPerception
from diffndm_b200.weights import DynamicsConfig, random_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = DynamicsConfig()
dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 1e-3 / 0.3 * 0.3), max_nodes=B * 800, max_edges=B * 40000, max_samples=B).eval()
px, pt = synthetic.synthetic_pocket(7, 330)
sizes = synthetic.synthetic_ligand_sizes(7, B)
onehot = np.eye(10, dtype=np.float32)[pt]
pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))),
          'size': torch.tensor([len(px)] * B), 'mask': torch.arange(B).repeat_interleave(len(px))}
smp = ConditionalSampler(dyn, timesteps=500)
torch.manual_seed(0)
t0 = time.time()
xh, xp, lm, pm = smp.sample_given_pocket(pocket, sizes, timesteps=500)
torch.cuda.synchronize()
dt = time.time() - t0
x = xh[:, :3]
com = torch.zeros((B, 3), device=x.device).index_add_(0, lm, x) / torch.bincount(lm)[:, None]
print(f'B={B} 500 steps in {dt:.2f} s ({B / dt:.1f} ligands/s incl. Python loop), finite={bool(torch.isfinite(xh).all())}, '
      f'|COM|max={float(com.abs().max()):.2e}, |x|max={float(x.abs().max()):.2f}, flags={dyn.engine.read_flags()}')
print('atom types', torch.bincount(xh[:, 3:].argmax(1), minlength=10).tolist())
info = {'bonds1': [[154.0] * 10] * 10, 'bonds2': [[134.0] * 10] * 10, 'bonds3': [[120.0] * 10] * 10}
st = BondPerception(dyn.engine, info)(x.contiguous(), xh[:, 3:].argmax(1), lm, B, return_matrices=False)
print('bonds per molecule', st['n_bonds'].tolist()[:10], 'fragments', st['n_components'].tolist()[:10])
