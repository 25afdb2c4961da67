"""clock64 milestones of CTA 0 (leader of pair 0) of the fused node kernel.  Needs a library built with
DNDM_EXTRA_NVCC_FLAGS=-DDNDM_EK_TRACE (development only).
MMA issuer: start | per row block: layer 1 committed, hid tile seen, layer 2 committed, h tile seen, each projection group committed
drain warp 2: start | per use: accumulator seen, accumulator released
producer: start | per row block: layer-1 loads issued, W4 loads issued, projection loads issued"""
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
buf = torch.zeros(384, dtype=torch.int64, device=dev)
dyn.engine.lib.dndm_debug_copy(dyn.engine._h, 7, ctypes.c_void_p(buf.data_ptr()), buf.numel() * 8, None)
torch.cuda.synchronize()
tr = buf.cpu().numpy()
base = min(x for x in tr if x > 0)
for name, lo in (('mma', 0), ('drain', 128), ('producer', 256)):
    v = [int(x - base) for x in tr[lo:lo + 128] if x > 0]
    print(name, v)
