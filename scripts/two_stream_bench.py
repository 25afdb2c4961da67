"""Experiment: the bench step as K independent sub-batches (B / K ligands each) on K streams, each with its own engine and
CUDA graph, so that one sub-batch's HBM-bound kernels (segment reduce, node GEMMs) can run beside another's XU-bound edge
kernel.  DNDM_SMS=<n> sizes every persistent grid for n SMs (the other SMs stay free for the other stream's kernels).

    python scripts/two_stream_bench.py --k 2 --steps 40          # prints ms per round (= per step of the whole batch)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from diffndm_b200 import engine as E  # noqa: E402
from diffndm_b200.sampler import ConditionalSampler  # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init  # noqa: E402

T = bench.T_STEPS


class SubBatch:
    def __init__(self, idx, B, dev, stream):
        self.B, self.stream = B, stream
        px, pt, sizes, b = bench.make_inputs(0, B)
        n_l, n_p = len(b['lig_mask']), len(b['pocket_mask'])
        N = n_l + n_p
        cfg = DynamicsConfig()
        with torch.cuda.stream(stream):
            dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=N + 256, max_edges=int(N * 40) + 4096,
                                     max_samples=B, check_nan=False)
            dyn.compute_pocket_output = False
            self.eng = eng = dyn.engine
            eng.set_static_masks(True)
            smp = ConditionalSampler(dyn, timesteps=T)
            gam = smp.gamma
            self.coef_tab = smp.step_coefficients(gam[:-1], gam[1:]).to(dev)
            self.t_tab = (torch.arange(1, T + 1, dtype=torch.float32) / T).to(dev)
            self.z = torch.from_numpy(b['xh_lig']).to(dev)
            self.xp = torch.from_numpy(b['xh_pocket']).to(dev)
            self.lig_mask = torch.from_numpy(b['lig_mask']).to(dev)
            self.pocket_mask = torch.from_numpy(b['pocket_mask']).to(dev)
            self.t_buf = torch.zeros(B, 1, device=dev)
            self.coef_buf = torch.zeros(B, 3, device=dev)
            self.eps = torch.zeros_like(self.z)
            self.noise = torch.zeros_like(self.z)
            self.isig_tab = (1.0 / smp.sigma_tab[1:]).to(dev)
            self.nais_tab = (-smp.alpha_tab[1:] / smp.sigma_tab[1:]).to(dev)
            self.score = bench.SyntheticScore(torch.from_numpy(b['x0_target']).to(dev), self.xp.clone(), self.lig_mask,
                                              n_p // B, B, dev)
            self.set_step(T - 1)
            self.body()
            stream.synchronize()
            self.body()
            stream.synchronize()
            eng.set_static_masks(True)
            self.graph = torch.cuda.CUDAGraph()
            if stream == torch.cuda.default_stream():
                with torch.cuda.graph(self.graph):
                    self.body()
            else:
                with torch.cuda.graph(self.graph, stream=stream):
                    self.body()
        stream.synchronize()

    def body(self):
        self.noise.normal_()
        self.eng.forward(self.z, self.xp, self.t_buf, self.lig_mask, self.pocket_mask, self.B, out_lig=self.eps, want_pocket=False)
        self.score.apply(self.eps, self.z, self.xp)
        self.eng.sampler_step(self.z, self.eps, self.noise, self.xp, self.coef_buf, self.lig_mask, self.pocket_mask, self.B,
                              z_out=self.z, pocket_out=self.xp, check_com=True)

    def set_step(self, s):
        self.t_buf.copy_(self.t_tab[s].expand(self.B, 1))
        self.coef_buf.copy_(self.coef_tab[s].expand(self.B, 3))
        self.score.set_step(self.isig_tab[s], self.nais_tab[s])

    def step(self, s, set_scalars=True):
        with torch.cuda.stream(self.stream):
            if set_scalars:
                self.set_step(s)
            self.graph.replay()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--k', type=int, default=2)
    ap.add_argument('--batch', type=int, default=100)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--default-stream', action='store_true')
    ap.add_argument('--frozen-scalars', action='store_true', help='timed loop replays the graph only (no per-step scalar copies)')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    assert args.batch % args.k == 0
    subs = [SubBatch(i, args.batch // args.k, dev, torch.cuda.current_stream() if args.default_stream else torch.cuda.Stream())
            for i in range(args.k)]
    s = T - 1
    s_hi = T // 2 + args.steps // 2
    while s > s_hi:
        for sb in subs:
            sb.step(s)
        s -= 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main_stream = torch.cuda.current_stream()
    e0.record(main_stream)
    for sb in subs:
        sb.stream.wait_event(e0)
    for _ in range(args.steps):
        for sb in subs:
            sb.step(s, not args.frozen_scalars)
        s -= 1
    for sb in subs:
        main_stream.wait_stream(sb.stream)
    e1.record(main_stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    stats = [sb.eng.graph_stats_full() for sb in subs]
    print(json.dumps({'k': args.k, 'sms': os.environ.get('DNDM_SMS'), 'batch': args.batch, 'frozen_scalars': args.frozen_scalars, 'ms_per_step': ms,
                      'ligands_per_s': args.batch / (bench.CALLS_PER_TRAJ * ms * 1e-3),
                      'edges': [int(x[0]) for x in stats]}))


if __name__ == '__main__':
    main()
