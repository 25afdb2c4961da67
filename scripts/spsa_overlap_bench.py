"""One SPSA guidance event (conditional_model.py:760-813: 2k = 20 perturbed copies of B ligands, x0 look-ahead, host
scores, update) with a host scorer of RDKit-like cost behind ``hostpool.PooledReward``: plain (denoise everything, then
score) against overlapped (score the +U half on the worker processes while the GPU denoises the -U half).
Usage: python scripts/spsa_overlap_bench.py [B=20] [pocket_atoms=330] [ms_per_molecule=0.3] [workers=8]
Prints one JSON object (median of 5 events per mode after a warm-up event)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, synthetic                          # noqa: E402
from diffndm_b200.hostpool import PooledReward                           # noqa: E402
from diffndm_b200.sampler import ConditionalSampler                      # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init             # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_p = int(sys.argv[2]) if len(sys.argv) > 2 else 330
MS = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
workers = int(sys.argv[4]) if len(sys.argv) > 4 else 8
os.environ['DNDM_SCORE_MS'] = str(MS)


def slow_score(x, types):
    """Stand-in for build_molecule + QED/SA: busy for DNDM_SCORE_MS milliseconds, then a geometric score."""
    t_end = time.perf_counter() + float(os.environ.get('DNDM_SCORE_MS', '0.3')) * 1e-3
    while time.perf_counter() < t_end:
        pass
    return -float(np.sqrt(((x - x.mean(0)) ** 2).sum(1).mean())) if len(x) else 0.0


def main():
    dev = torch.device('cuda', 0)
    cfg = DynamicsConfig()
    k = 10
    dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 1e-3), max_nodes=2 * k * B * (n_p + 50) + 1024,
                             max_edges=2 * k * B * (n_p + 50) * 24, max_samples=2 * k * B, check_nan=False).eval()
    smp = ConditionalSampler(dyn, timesteps=500)
    px, pt = synthetic.synthetic_pocket(7, n_p)
    b = synthetic.make_batch(px, pt, synthetic.synthetic_ligand_sizes(7, B), 7)
    tl = lambda a: torch.from_numpy(a).to(dev)
    args = (tl(b['xh_lig']), tl(b['xh_pocket']), tl(b['lig_mask']), tl(b['pocket_mask']), torch.full((B, 1), 20 / 500), B, 1e-3)
    out = {'batch': B, 'pocket_atoms': n_p, 'k': k, 'molecules_per_event': 2 * k * B, 'ms_per_molecule': MS, 'workers': workers}
    with PooledReward(slow_score, workers=workers) as pool:
        for name, overlap in (('plain', False), ('overlapped', True), ('plain_again', False)):
            smp.overlap_scoring = overlap
            times = []
            for rep in range(6):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                smp.my_update_z_lig(*args, pool, guidance_scale=1e-3, k=k)
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            out[name + '_ms_per_event'] = round(1e3 * float(np.median(times[1:])), 2)
    # the two legs alone: GPU part with a free scorer, host part on ready data
    smp.overlap_scoring = False
    free = lambda x, t, m: [0.0] * (int(m.max().item()) + 1)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        smp.my_update_z_lig(*args, free, guidance_scale=1e-3, k=k)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    out['gpu_only_ms_per_event'] = round(1e3 * float(np.median(ts[1:])), 2)
    out['host_only_ms_per_event_ideal'] = round(2 * k * B * MS / workers, 2)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
