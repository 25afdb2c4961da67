"""Per-tile timeline of the pair edge kernel (CTA 0 = leader of pair 0) from clock64 stamps.  Needs a library built with
DNDM_EXTRA_NVCC_FLAGS=-DDNDM_EK_TRACE (development only).

events: 0 producer warp 0 tile start | 1 metadata of the next tile published | 2 A[buf] free seen | 12 first half stored
        3 second half stored | 4 after the a_full arrive | 11 issuer: a_full seen | 5 issuer: accumulator free | 6 MMAs committed
        7 epilogue: MMA done seen | 8 epilogue: tile drained | 9 / 10 producer warp 15 start / stored
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
print('it | start meta A_free half0 half1 arrived | a_full_seen acc_free committed | epi_start epi_done | w15_start w15_stored')
for it in range(4, 28):
    r = tr[it] - base
    print(it, '|', r[0], r[1], r[2], r[12], r[3], r[4], '|', r[11], r[5], r[6], '|', r[7], r[8], '|', r[9], r[10])
s = slice(6, 28)
d = lambda a, b: (tr[s, a] - tr[s, b]).mean()
print('mean tile period (cycles):', np.diff(tr[6:29, 0]).mean())
print('producer: meta', d(1, 0), ' wait A_free', d(2, 1), ' half0', d(12, 2), ' half1', d(3, 12), ' fence+arrive', d(4, 3))
print('issuer: a_full seen after w0 arrive', d(11, 4), ' acc_free wait', d(5, 11), ' issue', d(6, 5))
print('epilogue: start after commit', d(7, 6), ' drain', d(8, 7))
