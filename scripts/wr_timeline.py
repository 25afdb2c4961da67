"""Per-row-block timeline of the merged projection GEMM (gemm_pair_kernel<256,256>, CTA 0 = leader of pair 0) from clock64 stamps.
Needs a library built with DNDM_EXTRA_NVCC_FLAGS=-DDNDM_EK_TRACE (development only).
events: 0 producer: row block start | 1 producer: 4 TMA loads issued | 2 MMA: first k-block landed | 3 MMA: last k-block
landed | 4 MMA: committed | 5 drain warp: accumulator ready | 6 drain warp: row block stored"""
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
buf = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
dyn.engine.lib.dndm_debug_copy(dyn.engine._h, 6, ctypes.c_void_p(buf.data_ptr()), buf.numel() * 8, None)
torch.cuda.synchronize()
tr = buf.cpu().numpy().reshape(64, 8)
base = tr[0, 0]
print('kernel entry 0 | set-up done', tr[2, 7] - tr[0, 7], '| predecessor done (pdl_wait)', tr[3, 7] - tr[0, 7], '| exit', tr[1, 7] - tr[0, 7])
base = tr[0, 7]
print('it | prod_start loads_issued | first_kb last_kb committed | acc_ready stored')
for it in range(0, 9):
    r = tr[it] - base
    print(it, '|', r[0], r[1], '|', r[2], r[3], r[4], '|', r[5], r[6])
