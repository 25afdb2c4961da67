"""Measurement scaffolding: per-tile timeline of CTA 0 of the block-2 GCL launch (DNDM_TIMELINE=1)."""
import os, sys, ctypes
os.environ['DNDM_TIMELINE'] = '1'
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, synthetic
from diffndm_b200.weights import DynamicsConfig, random_init
B = 100
px, pt = synthetic.synthetic_pocket(0); sizes = synthetic.synthetic_ligand_sizes(0, B); b = synthetic.make_batch(px, pt, sizes, 0)
N = len(b['lig_mask']) + len(b['pocket_mask'])
cfg = DynamicsConfig(); dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=N + 256, max_edges=N * 40, max_samples=B, check_nan=False)
dev = torch.device('cuda'); tl = lambda a: torch.from_numpy(a).to(dev)
args = (tl(b['xh_lig']), tl(b['xh_pocket']), torch.full((B, 1), 0.5, device=dev), tl(b['lig_mask']), tl(b['pocket_mask']))
for _ in range(3): dyn(*args, n_samples=B)
torch.cuda.synchronize()
buf = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
dyn.engine.lib.dndm_debug_copy(dyn.engine._h, 5, ctypes.c_void_p(buf.data_ptr()), 64 * 8 * 8, None)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(64, 8).astype(np.float64)
t0 = t[0, 0]
names = ['P.start', 'P.Afree', 'P.done', 'MMA.iss', 'E.start', 'E.pass1', 'E.end']
print('tile ' + ' '.join(f'{n:>9s}' for n in names) + '   (us since first stamp; producer = warp 8, epilogue = group it&1)')
for i in range(64):
    if t[i, 0] == 0: break
    print(f'{i:4d} ' + ' '.join(f'{(t[i, k] - t0) / 1.9e3:9.2f}' if t[i, k] else '        -' for k in range(7)))
d = t[:, 2] - t[:, 1]
print('P.done - P.Afree (us):', ' '.join(f'{x / 1.9e3:.1f}' for x in d[t[:, 0] > 0]))
