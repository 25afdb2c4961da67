"""Per-tile timeline of the GCL edge kernel (CTA 0) from clock64 stamps.  Needs a library built with
DNDM_EXTRA_NVCC_FLAGS=-DDNDM_EK_TRACE (development only).

events: 0 producer warp 0 tile start | 1 first half computed | 2 previous MMA done (A free) | 3 second half stored
        4 after producer barrier | 5 issuer: accumulator free | 6 issuer: MMAs committed | 7 epilogue: MMA done seen
        8 epilogue: tile drained | 9 producer warp 15 tile start | 10 producer warp 15 second half stored
"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, synthetic
from diffndm_b200.weights import DynamicsConfig, random_init

dev = torch.device('cuda')
cfg = DynamicsConfig()
dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3)).eval()
px, pt = synthetic.synthetic_pocket(0, 330)
b = synthetic.make_batch(px, pt, synthetic.synthetic_ligand_sizes(0, 100), 0)
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(b['xh_lig']), t(b['xh_pocket']), torch.full((100, 1), 0.5, device=dev), t(b['lig_mask']), t(b['pocket_mask']))
for _ in range(3):
    dyn(*args, n_samples=100)
torch.cuda.synchronize()
buf = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
n = dyn.engine.lib.dndm_debug_copy(dyn.engine._h, 5, ctypes.c_void_p(buf.data_ptr()), buf.numel() * 8, None)
torch.cuda.synchronize()
tr = buf.cpu().numpy().reshape(64, 16)
base = tr[4, 0]
print('it | tile_start half0_done A_free half1_stored after_bar | acc_free committed | epi_start epi_done | w15_start w15_stored')
for it in range(4, 28):
    r = tr[it] - base
    print(it, '|', r[0], r[1], r[2], r[3], r[4], '|', r[5], r[6], '|', r[7], r[8], '|', r[9], r[10])
per = np.diff(tr[6:30, 0]).mean()
print('mean tile period (cycles):', per)
print('mean half0 compute', (tr[6:30, 1] - tr[6:30, 0]).mean(), ' wait A_free', (tr[6:30, 2] - tr[6:30, 1]).mean(),
      ' half0 store + half1', (tr[6:30, 3] - tr[6:30, 2]).mean(), ' barrier wait', (tr[6:30, 4] - tr[6:30, 3]).mean(),
      ' acc_free wait', (tr[6:30, 5] - tr[6:30, 4]).mean(), ' issue', (tr[6:30, 6] - tr[6:30, 5]).mean())
print('epilogue: wait->start after commit', (tr[6:30, 7] - tr[6:30, 6]).mean(), ' drain', (tr[6:30, 8] - tr[6:30, 7]).mean())
print('warp15: start lag', (tr[6:30, 9] - tr[6:30, 0]).mean(), ' stored lag vs w0', (tr[6:30, 10] - tr[6:30, 3]).mean())
