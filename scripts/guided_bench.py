"""Throughput of the guided sampling modes (BASELINE configs 3-5 shapes) through the public sampler API, with a cheap
synthetic reward standing in for the host chemistry (RDKit QED/SA is external and not part of the path):
plain, SPSA (s <= 30, every 2nd step, k = 10 -> 2 batched denoiser calls on 20 x B samples per event), ATP (s <= 50,
every 10th step, 5 candidate groups) and SPSA + ATP, B = 20 ligands on one 330-atom synthetic pocket, 500 steps.
Prints one JSON object; four trajectories per mode, the median of the last three is reported (the first one warms the
allocator up; the rest still includes one CUDA-graph capture per trajectory and the Python loop)."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, synthetic
from diffndm_b200.sampler import ConditionalSampler
from diffndm_b200.weights import DynamicsConfig, random_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_p = int(sys.argv[2]) if len(sys.argv) > 2 else 330
cfg = DynamicsConfig()
fan = 20                                            # SPSA: 2k copies of the batch
dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=fan * B * (n_p + 50) + 1024,
                         max_edges=fan * B * (n_p + 50) * 24, max_samples=fan * B).eval()
px, pt = synthetic.synthetic_pocket(7, n_p)
sizes = synthetic.synthetic_ligand_sizes(7, B)
onehot = np.eye(10, dtype=np.float32)[pt]
pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))),
          'size': torch.tensor([len(px)] * B), 'mask': torch.arange(B).repeat_interleave(len(px))}
smp = ConditionalSampler(dyn, timesteps=500)
# stationary workload: the exact score of a point-mass pose in the pocket keeps the ligands there, as a
# trained model would; guidance copies of the batch reuse the pose rows
pose = synthetic.synthetic_ligand_pose(7, sizes, px.mean(axis=0, dtype=np.float64))
pose[:, :3] -= px[0]
smp.eps_transform = synthetic.PointMassScore(pose, smp.gamma, len(px), smp.T, torch.device('cuda'))


def reward(x, types, mask):                        # radius of gyration per molecule, one device reduction + one D2H
    n = int(mask.max().item()) + 1
    cnt = torch.bincount(mask, minlength=n).clamp(min=1).float()
    mean = torch.zeros((n, 3), device=x.device).index_add_(0, mask, x) / cnt[:, None]
    r2 = torch.zeros(n, device=x.device).index_add_(0, mask, ((x - mean[mask]) ** 2).sum(1)) / cnt
    return (-r2.sqrt()).tolist()


out = {'batch': B, 'pocket_atoms': n_p, 'timesteps': 500}
for name, kw in [('plain', {}), ('spsa', dict(spsa=1)), ('atp', dict(svdd=1)), ('spsa_atp', dict(spsa=1, svdd=1))]:
    dt = None
    err, times = None, []
    for rep in range(4):
        torch.manual_seed(rep)
        torch.cuda.synchronize()
        t0 = time.time()
        try:
            xh, xp, lm, pm = smp.sample_given_pocket(pocket, sizes, timesteps=500, reward_fn=reward, **kw)
        except (AssertionError, ValueError, RuntimeError) as ex:      # what test.py:164-168 of the reference catches and retries
            err = f'{type(ex).__name__}: {ex}'
            dyn.engine.set_static_masks(False)
            break
        torch.cuda.synchronize()
        times.append(time.time() - t0)
    dt = float(np.median(times[1:])) if len(times) > 1 else float('nan')      # the first trajectory warms the allocator up
    out[name] = {'error': err} if err else {'seconds_per_trajectory_batch': round(dt, 4), 'ligands_per_s': round(B / dt, 2),
                                            'all_seconds': [round(t, 3) for t in times],
                                            'finite': bool(torch.isfinite(xh).all()), 'flags': dyn.engine.read_flags()}
print(json.dumps(out))
