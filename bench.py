#!/usr/bin/env python
"""Benchmark of the hot path: ligands/sec for 500-step fullatom_cond conditional sampling (BASELINE.json).

A *step* is one pass of the hot path over one batch: one reverse-diffusion step = EGNN denoiser forward
(radius graph + 6 equivariant blocks) + fused p(z_s|z_t) update, for one synthetic CrossDocked-shaped pocket
replicated over B ligands (configs[1] of BASELINE.json: 100 ligands per pocket, pockets sharded over GPUs; rank r
works on pocket r).  A 500-step trajectory makes 501 denoiser calls, so

    value [ligands/s] = n_gpus * B / (501 * seconds_per_step)

`value` is measured with the state resident in HBM and the step captured in a CUDA graph; `e2e` runs the same step
through the public Python API with HOST (pinned) buffers, copying the step inputs H2D and the result D2H inside the
timed region.  `roofline` times the dominant kernel (fused GCL edge kernel) with CUDA events in an eager pass over
the same inputs.  `cpu_baseline` / `--impl reference` time the numpy oracle (a restatement of the reference's PyTorch
CPU path; the reference itself cannot travel to the GPU box) on the host cores on a bounded sample.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout is reserved for the ONE JSON line.  NCCL (version banner), torch and child processes write to file descriptor 1
# directly, so the descriptor itself is pointed at stderr for the whole run and the JSON line goes to a private duplicate
# of the original stdout.
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
sys.stdout.flush()
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit_json(line):
    os.write(_JSON_FD, (json.dumps(line) + '\n').encode())


T_STEPS = 500
CALLS_PER_TRAJ = T_STEPS + 1
EXEC_FLOP_PER_EDGE_GCL = 2 * 256 * 256                     # second edge-MLP layer, the GEMM the kernel runs
REF_FLOP_PER_EDGE_GCL = 2 * 514 * 256 + 2 * 256 * 256 + 512  # reference formulation of the same MLP


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=100, help='ligands per pocket (BASELINE configs[1]: 100)')
    ap.add_argument('--cpu-batch', type=int, default=None, help='ligands in the bounded CPU sample (default: one per host core, 8..32)')
    ap.add_argument('--pocket-atoms', type=int, default=None, help='override the pocket size (default: 330, the distribution mean)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-job', action='store_true', help='skip the whole-job section (pocket queue -> sampling -> SDF files; ATP over the ranks)')
    ap.add_argument('--job-pockets', type=int, default=4, help='pockets per GPU in the job section (weak scaling: 4 N pockets on N GPUs)')
    args = ap.parse_args()
    if args.cpu_batch is None:
        args.cpu_batch = max(8, min(32, os.cpu_count() or 8))
    return args


MIN_LIG_RECV_SHARE = 0.11      # ligand-receiver edges / edges below this: the ligands have left the pocket
MIN_LAST_BLOCK_SHARE = 0.27    # last-block edges / edges
REFERENCE_BUDGET_S = 150.0     # wall-clock bound of the `--impl reference` arm
POCKET_ATOMS = 330     # mean of the CrossDocked-shaped pocket-size distribution; same size on every rank so
                       # that the weak-scaling runs compare equal per-GPU work (geometry / ligand sizes differ per rank)


def make_inputs(rank, batch):
    """Pocket, ligand sizes, the z_T batch (make_batch) and the synthetic data point x_0 in the batch's frame.

    Random-init weights cannot denoise: with them alone a reverse trajectory inflates by 1/alpha_T ~ 45x and the ligands
    leave the pocket within ~25 steps (no ligand-pocket edges left -- a degenerate, too cheap workload).  The bench therefore
    adds the exact score of a point-mass data distribution to the network output, eps = eps_net + (z_t - alpha_t x_0) /
    sigma_t, with x_0 a ligand-shaped point cloud in the pocket cavity (synthetic.synthetic_ligand_pose): every step is
    then a real p(z_s | z_t) move of a trajectory that converges to x_0 like a trained model's does to a molecule, the
    ligands stay in the pocket from z_T to z_0, and the edge counts are stationary (emitted and asserted below)."""
    from diffndm_b200 import synthetic
    px, pt = synthetic.synthetic_pocket(rank, POCKET_ATOMS)
    sizes = synthetic.synthetic_ligand_sizes(rank, batch)
    b = synthetic.make_batch(px, pt, sizes, rank)
    n_p = len(px)
    com = px.mean(axis=0, dtype=np.float64)
    pose = synthetic.synthetic_ligand_pose(rank, sizes, com)
    # make_batch moved every sample into its ligand-COM-free frame: apply the same per-sample translation to x_0
    shift = b['xh_pocket'][::n_p, :3] - px[0][None, :]                    # [B,3]
    pose[:, :3] += shift[b['lig_mask']]
    b['x0_target'] = pose
    return px, pt, sizes, b


class SyntheticScore:
    """eps += (z - alpha_t * x_0(frame)) / sigma_t, where x_0 follows the pocket's accumulated translation (the pocket is
    rigid: its first atom gives the translation).  Six small torch kernels on [N_l, 13] rows inside the timed region -- work
    the real path does not have, so the measured step is slightly pessimistic."""

    def __init__(self, x0_target, xh_pocket0, lig_mask, n_p, B, dev):
        import torch
        self.x0 = x0_target
        self.first = torch.arange(B, device=dev) * n_p
        self.p0 = xh_pocket0[self.first].clone()
        self.lig_mask = lig_mask
        self.isig = torch.zeros(1, 1, device=dev)          # 1 / sigma_t
        self.nais = torch.zeros(1, 1, device=dev)          # -alpha_t / sigma_t

    def set_step(self, isig_t, nais_t):
        self.isig.copy_(isig_t)
        self.nais.copy_(nais_t)

    def apply(self, eps, z, xp):
        tgt = self.x0 + (xp[self.first] - self.p0)[self.lig_mask]
        eps.addcmul_(z, self.isig)
        eps.addcmul_(tgt, self.nais)


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_step_seconds(batch, n_steps, warmup, rank=0):
    """Seconds per denoising step of the numpy port on ``batch`` ligands, using every host core: the samples of a batch
    are independent (block-diagonal graph), so the batch is cut into one group of samples per core and the groups run on
    a thread pool (numpy's element-wise kernels are single-threaded; matmuls and ufuncs release the GIL), BLAS limited to
    one thread per group.  Measured in the build container (8 cores, 8 ligands, 330-atom pocket): 2.04 s per step against
    2.32 s for the reference's own torch CPU forward on the same cores (the serial numpy port took 7.2 s), DESIGN.md section 5."""
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    groups = max(1, min(cores, batch))
    try:   # torchrun exports OMP_NUM_THREADS=1; with one sample group per core BLAS stays single-threaded per group
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=max(1, cores // groups))
    except Exception:
        pass
    from oracle import egnn_oracle as O
    from diffndm_b200.weights import DynamicsConfig, random_init
    W = random_init(DynamicsConfig(), 0, 0.3)
    px, pt, sizes, b = make_inputs(rank, batch)
    cfg = O.OracleConfig()
    g = O.gamma_table(T_STEPS, cfg.noise_precision, 2.0)
    rng = np.random.default_rng(0)
    # per group: its samples' rows, masks renumbered from 0
    bounds = np.linspace(0, batch, groups + 1).astype(int)
    parts = []
    for gi in range(groups):
        lo, hi = bounds[gi], bounds[gi + 1]
        sl = (b['lig_mask'] >= lo) & (b['lig_mask'] < hi)
        sp = (b['pocket_mask'] >= lo) & (b['pocket_mask'] < hi)
        parts.append({'z': b['xh_lig'][sl], 'xp': b['xh_pocket'][sp], 'lm': b['lig_mask'][sl] - lo, 'pm': b['pocket_mask'][sp] - lo,
                      'n': int(hi - lo), 'seed': gi, 'x0': b['x0_target'][sl],
                      'first': np.arange(int(hi - lo)) * b['n_pocket'], 'p0': b['xh_pocket'][sp][::b['n_pocket']].copy()})
    alpha_tab, sigma_tab = np.sqrt(O.sigmoid(-g)), np.sqrt(O.sigmoid(g))
    # start where the GPU arm's timed window sits: the middle of the trajectory, z_t ~ q(z_t | x_0)
    s_start = T_STEPS // 2
    for part in parts:
        nz = np.random.default_rng(77 + part['seed']).standard_normal(part['z'].shape).astype(np.float32)
        part['z'], part['xp'] = O.noised_representation(part['x0'], part['xp'], nz, np.full(part['n'], g[s_start + 1]),
                                                        part['lm'], part['pm'])

    def one(part, s):
        n = part['n']
        tt = np.full((n, 1), (s + 1) / T_STEPS, np.float32)
        eps, _ = O.dynamics_forward(W, part['z'], part['xp'], tt, part['lm'], part['pm'], cfg)
        tgt = part['x0'] + (part['xp'][part['first']] - part['p0'])[part['lm']]   # synthetic score, as in the GPU arm (make_inputs)
        eps = (eps + (part['z'] - alpha_tab[s + 1] * tgt) / sigma_tab[s + 1]).astype(np.float32)
        noise = np.random.default_rng(1000 * s + part['seed']).standard_normal(part['z'].shape).astype(np.float32)
        part['z'], part['xp'] = O.sample_p_zs_given_zt(part['z'], part['xp'], eps, noise, np.full(n, g[s]), np.full(n, g[s + 1]),
                                                       part['lm'], part['pm'])

    times = []
    s = s_start
    with ThreadPoolExecutor(max_workers=groups) as pool:
        for i in range(warmup + n_steps):
            t0 = time.perf_counter()
            list(pool.map(lambda p: one(p, s), parts))
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            s = max(s - 1, 0)
    return float(np.mean(times)), len(b['pocket_mask']) // batch


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cores = os.cpu_count()
    # bounded sample: the whole --steps K --warmup W run has to end within a few minutes, so the number of ligands per CPU
    # step shrinks towards one per core until K steps fit REFERENCE_BUDGET_S (a step then costs one ligand's forward)
    cpu_batch, warm = args.cpu_batch, min(args.warmup, 2)
    probe, _ = cpu_step_seconds(cpu_batch, 1, 0)
    total = probe * (args.steps + warm)
    if total > REFERENCE_BUDGET_S:      # one ligand per core is the floor: fewer ligands than cores do not shorten a step
        cpu_batch = max(min(cores, cpu_batch), int(cpu_batch * REFERENCE_BUDGET_S / total))
    sec, n_p = cpu_step_seconds(cpu_batch, args.steps, warm)
    val = cpu_batch / (CALLS_PER_TRAJ * sec)
    sample = (f'{args.steps} denoising steps (numpy oracle forward + p(z_s|z_t)) on {cpu_batch} ligands of the same '
              f'synthetic pocket ({n_p} atoms), one group of samples per core on a thread pool; per-step cost scaled to 501 '
              f'calls per trajectory')
    line = {
        'impl': 'reference', 'metric': 'ligands/sec (500-step fullatom_cond sampling)', 'value': val, 'unit': 'ligands/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.batch),
        'cpu_baseline': {'value': val, 'unit': 'ligands/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'ligands/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit_json(line)


def workload_config(batch):
    return {'workload': f'crossdocked_fullatom_cond conditional sampling, synthetic CrossDocked-shaped pockets x {batch} '
                        f'ligands, {T_STEPS} steps (BASELINE configs[1]); one {POCKET_ATOMS}-atom pocket per GPU, random-init weights',
            'ligands_per_pocket': batch, 'timesteps': T_STEPS,
            'l2': 'per-step working set (node projections + activations) exceeds the 126 MB L2; no explicit flush'}


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 10 ms from a thread (the timed
    region can be shorter than nvidia-smi's start-up), with the nvidia-smi loop of the profiling recipe as fallback."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.samples = []          # (time, sm_mhz, set(reasons))
        self.power = []            # (time, board power in W)
        self.power_limit = None
        self.smax = None
        self.stop_flag = False
        self.proc = None
        self.mode = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = gpu_index
            if vis and all(v.strip().isdigit() for v in vis.split(',')):
                idx = int(vis.split(',')[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            try:
                self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
            except Exception:
                pass
            self.mode = 'nvml'
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20',
                                          '-i', str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = 'smi'
            self.th = threading.Thread(target=self._read_smi, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        bits = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(get_reasons(self.h))
                self.samples.append((time.time(), mhz, {n for n, b in bits.items() if r & b}))
                try:            # board power next to the clock: a clock below max at ~the power limit explains itself
                    self.power.append((time.time(), nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
                except Exception:
                    pass
            except Exception:
                pass
            time.sleep(0.01)

    def _read_smi(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for l in self.proc.stdout:
            p = [x.strip() for x in l.strip().split(',')]
            if len(p) < 7:
                continue
            try:
                self.smax = float(p[1])
                self.samples.append((time.time(), float(p[0]),
                                     {n for n, v in zip(names, p[3:7]) if v.lower().startswith('active')}))
            except ValueError:
                continue

    def stop(self, t0, t1):
        if self.mode is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml and nvidia-smi unavailable'], 'samples': 0}
        time.sleep(0.05)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, reasons = [], set()
        for ts, mhz, rs in self.samples:
            if t0 <= ts <= t1:
                sm.append(mhz)
                reasons |= rs
        out = {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.smax, 'reasons': sorted(reasons),
               'samples': len(sm), 'source': self.mode}
        pw = [w for ts, w in self.power if t0 <= ts <= t1]
        if pw:                     # NVML's power reading is a ~1 s running average: it lags a 0.2 s window
            out.update({'power_w_max': max(pw), 'power_limit_w': self.power_limit})
        return out


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from diffndm_b200 import engine as E
    from diffndm_b200.sampler import ConditionalSampler
    from diffndm_b200.weights import DynamicsConfig, random_init

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    B = args.batch
    px, pt, sizes, b = make_inputs(rank, B)
    n_l, n_p = len(b['lig_mask']), len(b['pocket_mask'])
    N = n_l + n_p
    cfg = DynamicsConfig()
    dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=N + 256, max_edges=int(N * 40) + 4096,
                             max_samples=B, check_nan=False)
    dyn.compute_pocket_output = False            # every conditional caller discards it (`eps, _ =`)
    eng = dyn.engine
    eng.set_static_masks(True)      # the masks of a trajectory are fixed tensors (as in ConditionalSampler.sample_given_pocket)
    smp = ConditionalSampler(dyn, timesteps=T_STEPS)

    # per-step scalars for every s (device tables): t, coefficients
    gam = smp.gamma
    coef_tab = smp.step_coefficients(gam[:-1], gam[1:]).to(dev)                 # [T,3]  step s -> s+1
    t_tab = (torch.arange(1, T_STEPS + 1, dtype=torch.float32) / T_STEPS).to(dev)  # t of step s

    z0 = torch.from_numpy(b['xh_lig']).to(dev)
    p0 = torch.from_numpy(b['xh_pocket']).to(dev)
    lig_mask = torch.from_numpy(b['lig_mask']).to(dev)
    pocket_mask = torch.from_numpy(b['pocket_mask']).to(dev)
    z, xp = z0.clone(), p0.clone()
    # per-step scalars as in the sampler's graphed reverse step (sampler._GraphedReverseStep): row s of `table` = (t, step
    # coefficients, the two scalars of the synthetic score) of step s; the captured step gathers row `s_dev` and decrements it
    Bp = (B + 3) // 4 * 4
    packed = torch.zeros(1, Bp + 3 * B + 4, device=dev)
    t_buf = packed[0, :B].view(B, 1)
    coef_buf = packed[0, Bp:Bp + 3 * B].view(B, 3)
    s_dev = torch.zeros(1, dtype=torch.long, device=dev)
    eps = torch.zeros_like(z)
    noise = torch.zeros_like(z)
    # synthetic score (see make_inputs): 1/sigma_t and -alpha_t/sigma_t of step s -> s+1's t
    isig_tab = (1.0 / smp.sigma_tab[1:]).to(dev)
    nais_tab = (-smp.alpha_tab[1:] / smp.sigma_tab[1:]).to(dev)
    score = SyntheticScore(torch.from_numpy(b['x0_target']).to(dev), p0, lig_mask, n_p // B, B, dev)
    score.isig = packed[:, Bp + 3 * B:Bp + 3 * B + 1]
    score.nais = packed[:, Bp + 3 * B + 1:Bp + 3 * B + 2]
    table = torch.zeros(T_STEPS, packed.shape[1], device=dev)
    table[:, :B] = t_tab[:, None]
    table[:, Bp:Bp + 3 * B] = coef_tab[:, None, :].expand(T_STEPS, B, 3).reshape(T_STEPS, 3 * B)
    table[:, Bp + 3 * B] = isig_tab
    table[:, Bp + 3 * B + 1] = nais_tab
    next_s = [None]

    def body():
        torch.index_select(table, 0, s_dev, out=packed)
        s_dev.sub_(1)
        noise.normal_()
        eng.forward(z, xp, t_buf, lig_mask, pocket_mask, B, out_lig=eps, want_pocket=False)
        score.apply(eps, z, xp)
        eng.sampler_step(z, eps, noise, xp, coef_buf, lig_mask, pocket_mask, B, z_out=z, pocket_out=xp, check_com=True)

    def set_step(s):                 # a launch only when s is not the successor of the previous step
        if next_s[0] != s:
            s_dev.fill_(s)
        next_s[0] = s - 1

    graph = None
    l0 = E.launch_count()
    set_step(T_STEPS - 1)
    body()                                   # eager warm-up (also sizes torch's caches)
    torch.cuda.synchronize()
    launches_per_step = E.launch_count() - l0
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        eng.set_static_masks(True)          # cache reset: the captured step derives the per-sample offsets itself, exactly
        graph = torch.cuda.CUDAGraph()      # like the sampler's graphed reverse step (sampler._GraphedReverseStep)
        with torch.cuda.graph(graph):
            body()
        next_s[0] = None                    # the dry runs moved the device-side step index

    def step(s):
        set_step(s)
        if graph is not None:
            graph.replay()
        else:
            body()

    def reset():
        z.copy_(z0)
        xp.copy_(p0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the timed window is centred on the trajectory's midpoint: the reverse steps from z_T down to its upper end run
    #      untimed (real steps, same graph), then W warm-up steps, then EXACTLY K timed steps s_hi, s_hi - 1, ...  With
    #      K >= T the whole trajectory is timed from z_T. ----
    reset()
    n_warm = max(args.warmup, 3)
    s = T_STEPS - 1
    s_hi = min(T_STEPS - 1, T_STEPS // 2 + args.steps // 2)
    while s > s_hi + n_warm:
        step(s)
        s -= 1
    for _ in range(n_warm):
        step(s)
        s = max(s - 1, 0)
    window = [int(s), int(max(s - args.steps + 1, 0))]
    barrier()
    clocks = ClockSampler(local)
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step(s)
        s = max(s - 1, 0)
    e1.record()
    barrier()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop(w0, w1)
    flags = eng.read_flags()
    if flags & (E.FLAG_EDGE_OVERFLOW | E.FLAG_NAN):
        raise RuntimeError(f'engine flags {flags} during the timed region')
    E_edges, E_lig, E_levels = eng.graph_stats_pruned()         # edges of the trailing (pruned) blocks, last block first
    E_last = E_levels[0]
    # the workload is only valid while the ligands sit in the pocket: ligand-receiver edges (ligand-ligand + ligand<-pocket)
    # and the last block's edge list (ligand receivers + their pocket senders) as shares of all edges.  The real 3rfm
    # complex (286 pocket atoms, 23-atom ligands) has 0.13 / ~0.2-0.3; ligands that left the pocket give 0.09 / 0.10.
    if E_lig < MIN_LIG_RECV_SHARE * E_edges or E_last < MIN_LAST_BLOCK_SHARE * E_edges:
        raise RuntimeError(f'degenerate bench state: ligand-receiver edges {E_lig}/{E_edges}, last-block edges {E_last}/{E_edges}')
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = tmax.item() / args.steps
    value = world * B / (CALLS_PER_TRAJ * ms_per_step * 1e-3)

    s_after, z_after, xp_after = int(s), z.clone(), xp.clone()      # where the e2e loop continues the trajectory

    # ---- roofline: the dominant kernel, CUDA events around every launch, eager pass over the same state ----
    roof = None
    prof = None
    if rank == 0:
        eng.set_profile(True)
        s2 = max(s, 0)
        n_prof = min(args.steps, 20)
        for _ in range(n_prof):
            set_step(s2)
            body()
            s2 = max(s2 - 1, 0)
        prof = eng.get_profile()
        eng.set_profile(False)
        g_ms, g_n = prof['gcl_edge_kernel']
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        # each launch is timed on its own by CUDA events in a ~50 ms eager pass (no power-cap samples at full clocks): the
        # burst figure is the applicable denominator (the sustained one belongs to seconds-long loops)
        peak = peaks.get('bf16_tflops')
        peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed per launch)'
        if peak is None:
            peak, peak_src = 1590.0, 'fallback (B200_PROFILING.md)'
        # six launches per forward: the trailing ones over the pruned edge lists E_levels, the others over all E edges
        # (exact dead-work elimination); achieved = executed FLOP of all timed launches / their total time
        n_layers = cfg.n_layers
        edges_timed = (g_n // n_layers) * ((n_layers - len(E_levels)) * E_edges + sum(E_levels))
        t_launch = g_ms / max(g_n, 1) * 1e-3
        achieved = edges_timed * EXEC_FLOP_PER_EDGE_GCL / (g_ms * 1e-3) / 1e12
        # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of ONE `ncu --set full` capture of this
        # kernel in this build, stored per edge by scripts/ncu_summary.py next to the raw pages under profiles/; a capture of
        # another build (other kernel name / version tag) is not used
        traffic, traffic_src = None, None
        try:
            cap = json.load(open(os.path.join(ROOT, 'profiles', 'gcl_traffic.json')))
            if cap.get('kernel_version') == E.load_library().dndm_version().decode():
                traffic = cap['dram_bytes_per_edge'] * (edges_timed / max(g_n, 1))
                traffic_src = cap.get('source')
        except Exception:
            pass
        roof = {'bound': 'tensor', 'kernel': 'edge_pair_kernel<GCL>', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': peak_src, 'limiter': 'XU pipe (MUFU.TANH for two SiLU per edge and channel + F2FP packs) and issue slots: see DESIGN.md section 5',
                'us_per_launch': t_launch * 1e6, 'edges_per_launch': edges_timed / max(g_n, 1), 'edges': E_edges,
                'edges_last_block': E_last, 'edges_trailing_blocks': E_levels,
                'flop_per_edge_executed': EXEC_FLOP_PER_EDGE_GCL,
                'achieved_reference_equivalent': (g_n // n_layers) * n_layers * E_edges * REF_FLOP_PER_EDGE_GCL / (g_ms * 1e-3) / 1e12,
                'step_share_ms': {k: v[0] / n_prof for k, v in prof.items()}}

    # ---- e2e: same step through the public API with HOST buffers (H2D of the step inputs, D2H of the result) ----
    e2e = None
    if not args.no_e2e:
        # Host side of the call: ONE pinned block holds every input of the step (state, pocket, time, step coefficients, both
        # masks) and ONE holds the outputs (new state, translated pocket), so a step costs one H2D and one D2H transfer
        # instead of six + two small ones (each ~10 us of latency on top of its bytes); the tensors handed to the public API
        # are views into the device mirrors of those blocks.
        def carve(spec, device, pin):
            sizes = [(-(-int(np.prod(shape)) * torch.empty((), dtype=dt).element_size() // 16)) * 16 for shape, dt in spec]
            block = torch.empty(sum(sizes), dtype=torch.uint8, device=device, pin_memory=pin)
            views, off = [], 0
            for (shape, dt), nbytes in zip(spec, sizes):
                n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
                views.append(block[off:off + n].view(dt).view(shape))
                off += nbytes
            return block, views
        in_spec = [(tuple(z0.shape), torch.float32), (tuple(p0.shape), torch.float32), ((B, 1), torch.float32),
                   ((B, 3), torch.float32), ((n_l,), torch.int64), ((n_p,), torch.int64), ((1, 2), torch.float32)]
        out_spec = [(tuple(z0.shape), torch.float32), (tuple(p0.shape), torch.float32)]
        # two identical pinned input blocks used in turn: the outputs of a step (new state, translated pocket) land in the
        # leading region of the OTHER block, which is the next step's input -- the trajectory goes through host memory
        # every step without a host-side memcpy
        hb = [carve(in_spec, 'cpu', True) for _ in range(2)]
        d_in, (dz, dp, dt_, dc, dml, dmp, dsc) = carve(in_spec, dev, False)
        d_out, (dzo, dpo) = carve(out_spec, dev, False)
        z_h, xp_h = z_after.cpu(), xp_after.cpu()
        for blk, (hz, hp, ht, hc, hm_l, hm_p, hsc) in hb:
            hz.copy_(z_h); hp.copy_(xp_h); hm_l.copy_(lig_mask.cpu()); hm_p.copy_(pocket_mask.cpu())
        coef_cpu, t_cpu = coef_tab.cpu(), t_tab.cpu()
        sc_cpu = torch.stack([isig_tab.cpu(), nais_tab.cpu()], dim=1)              # [T,2]: 1/sigma_t, -alpha_t/sigma_t
        score2 = SyntheticScore(score.x0, p0, dml, n_p // B, B, dev)
        score2.isig, score2.nais = dsc[:, 0:1], dsc[:, 1:2]                        # arrive with the step's inputs
        n_e2e = min(args.steps, 50)

        # One step through the public API with HOST buffers: H2D of every input of the call, denoiser forward, noise draw,
        # p(z_s|z_t), D2H of the new state and pocket.  The chain is sequential (the output of a step is the next step's
        # input, via the host).  Like the device-resident loop, the call sequence is captured in CUDA graphs, one per
        # direction of the block pair (copy nodes read / write the pinned host buffers at replay time).
        def e2e_body(i):
            d_in.copy_(hb[i][0], non_blocking=True)
            e_, _ = dyn(dz, dp, dt_, dml, dmp, n_samples=B)
            score2.apply(e_, dz, dp)
            nz = torch.randn_like(dz)
            eng.sampler_step(dz, e_, nz, dp, dc, dml, dmp, B, z_out=dzo, pocket_out=dpo, check_com=True)
            hb[1 - i][0][:d_out.numel()].copy_(d_out, non_blocking=True)

        e2e_graphs = None
        eng.set_static_masks(False)                 # the masks arrive from the host on every call
        if not args.no_graph:
            for blk, (hz, hp, ht, hc, hm_l, hm_p, hsc) in hb:
                ht.copy_(t_cpu[s_after].expand(B, 1))
                hc.copy_(coef_cpu[s_after].expand(B, 3))
                hsc.copy_(sc_cpu[s_after].reshape(1, 2))
            side2 = torch.cuda.Stream()
            side2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side2):
                e2e_body(0)
            torch.cuda.current_stream().wait_stream(side2)
            torch.cuda.synchronize()
            e2e_graphs = []
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    e2e_body(i)
                e2e_graphs.append(g)
            torch.cuda.synchronize()
            for blk, (hz, hp, ht, hc, hm_l, hm_p, hsc) in hb:      # the dry runs advanced the state: start again
                hz.copy_(z_h); hp.copy_(xp_h)
        turn = [0]

        def e2e_step(s):
            i = turn[0]
            hb[i][1][2].copy_(t_cpu[s].expand(B, 1))
            hb[i][1][3].copy_(coef_cpu[s].expand(B, 3))
            hb[i][1][6].copy_(sc_cpu[s].reshape(1, 2))
            if e2e_graphs is not None:
                e2e_graphs[i].replay()
            else:
                e2e_body(i)
            torch.cuda.synchronize()
            turn[0] = 1 - i

        s3 = s_after
        for _ in range(3):
            e2e_step(s3)
            s3 = max(s3 - 1, 0)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(n_e2e):
            e2e_step(s3)
            s3 = max(s3 - 1, 0)
        a1.record()
        barrier()
        t_e2e = torch.tensor([a0.elapsed_time(a1)], device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        ms_e2e = t_e2e.item() / n_e2e
        h2d = d_in.numel()
        d2h = d_out.numel()
        e2e = {'value': world * B / (CALLS_PER_TRAJ * ms_e2e * 1e-3), 'unit': 'ligands/s', 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': ms_e2e, 'steps': n_e2e, 'cuda_graph': e2e_graphs is not None,
               'transfers_per_step': 'one H2D (all inputs of the call in one pinned block) + one D2H'}

    # ---- the job: what the reference's my_test.py:68-90 does -- many pockets x B ligands, 500 steps each, one SDF per pocket --
    #      through the shared PocketQueue (largest pocket first), wall clock with a barrier on both sides, max over ranks.  Pocket
    #      sizes ~ clip(N(330, 80), 150, 700); job-pockets per GPU (weak scaling).  Plus one ATP trajectory (B = 20, 5 candidate
    #      groups) whose groups are split over the ranks: the NCCL all-gather of the path, one per event. ----
    job = None
    if not args.no_job:
        job = run_job_section(args, dyn, smp, world, rank, dev, B)

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, _ = cpu_step_seconds(args.cpu_batch, 6, 1)
        cpu = {'value': args.cpu_batch / (CALLS_PER_TRAJ * sec), 'unit': 'ligands/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': f'6 denoising steps of the numpy oracle (one group of samples per core) on {args.cpu_batch} ligands of the rank-0 pocket '
                         f'({n_p // B} atoms), {sec:.2f} s/step, scaled to 501 calls per trajectory'}

    if rank == 0:
        cfgd = workload_config(B)
        cfgd.update({'pocket_atoms_rank0': n_p // B, 'ligand_atoms_rank0': n_l, 'nodes': N, 'edges': E_edges,
                     'ligand_receiver_edges': E_lig, 'edges_last_block': E_last, 'edges_trailing_blocks': E_levels,
                     'ligand_receiver_share': E_lig / E_edges, 'last_block_share': E_last / E_edges,
                     'trajectory_window_s': window, 'cuda_graph': graph is not None,
                     'score': 'eps_net + (z_t - alpha_t x_0)/sigma_t, x_0 = synthetic ligand pose in the pocket (random-init '
                              'weights cannot denoise; see make_inputs)'})
        line = {
            'metric': 'ligands/sec (500-step fullatom_cond sampling)', 'value': value, 'unit': 'ligands/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': cfgd, 'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches_per_step * args.steps),
            'gpu_launches_per_step': int(launches_per_step), 'roofline': roof, 'cpu_baseline': cpu, 'job': job,
        }
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()


def run_job_section(args, dyn, smp, world, rank, dev, B):
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from diffndm_b200 import engine as E, synthetic
    from diffndm_b200.chem import BondPerception
    from diffndm_b200.datasets import crossdock_dataset_info
    from diffndm_b200.hostpool import PooledReward, radius_of_gyration_score
    from diffndm_b200.job import run_pocket_job, synthetic_pocket_sizes
    from diffndm_b200.sampler import ConditionalSampler
    from diffndm_b200.weights import DynamicsConfig, random_init
    P = args.job_pockets * world
    n_atoms = synthetic_pocket_sizes(P)
    nmax = int(n_atoms.max())
    cfg = DynamicsConfig()
    info = crossdock_dataset_info()
    big = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 0.3), max_nodes=B * (nmax + 60) + 1024, max_edges=B * (nmax + 60) * 48,
                             max_samples=max(B, 128), check_nan=False).eval()
    smp_job = ConditionalSampler(big, timesteps=T_STEPS)
    out_dir = tempfile.mkdtemp(prefix='dndm_job_')
    dt, gathered, idle = run_pocket_job(smp_job, BondPerception(big.engine, info), info, n_atoms, B, T_STEPS, out_dir)
    shutil.rmtree(out_dir, ignore_errors=True)
    done = sorted((p for g in gathered for p in g), key=lambda p: p['id'])
    assert [p['id'] for p in done] == list(range(P)), 'every pocket exactly once'
    job = {'pockets': P, 'ligands_per_pocket': B, 'timesteps': T_STEPS, 'seconds': dt, 'value': P * B / dt, 'unit': 'ligands/s',
           'pockets_per_rank': [len(g) for g in gathered], 'tail_idle_s_per_rank': idle,
           'pocket_atoms': [p['atoms'] for p in done], 'pocket_seconds': [p['s'] for p in done],
           'sample_seconds': [p['sample_s'] for p in done], 'output_seconds': [p['output_s'] for p in done],
           'setup_seconds': [p['setup_s'] for p in done],
           'ligand_receiver_share': [p['ligand_receiver_share'] for p in done],
           'what': 'PocketQueue -> pocket batch -> 500-step sampling (graph replay, captured per batch shape) -> GPU bond perception '
                   '-> largest fragment -> one SDF file per pocket; wall clock, barrier both sides, max over ranks'}
    # one ATP trajectory shared by all ranks: candidate groups g = 0..4 live on rank g % N, one all-gather per event
    Ba = 20
    px, pt = synthetic.synthetic_pocket(9000, POCKET_ATOMS)
    sizes = synthetic.synthetic_ligand_sizes(9000, Ba)
    onehot = np.eye(10, dtype=np.float32)[pt]
    pocket = {'x': torch.from_numpy(px).to(dev).repeat(Ba, 1), 'one_hot': torch.from_numpy(onehot).to(dev).repeat(Ba, 1),
              'size': torch.tensor([len(px)] * Ba, device=dev), 'mask': torch.arange(Ba, device=dev).repeat_interleave(len(px))}
    pose = synthetic.synthetic_ligand_pose(9000, sizes, px.mean(axis=0, dtype=np.float64))
    pose[:, :3] -= px[0]
    smp_job.eps_transform = synthetic.PointMassScore(pose, smp_job.gamma, len(px), smp_job.T, dev)
    reward = PooledReward(radius_of_gyration_score, workers=0)
    if world > 1:
        smp_job.set_distributed_atp(dist.group.WORLD, shared_seed=4242)
    else:
        torch.manual_seed(4242)
    times = []
    for rep in range(2):                      # the second run replays the captured reverse-step graphs
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        xh, _, lm, _ = smp_job.sample_given_pocket(pocket, sizes, timesteps=T_STEPS, svdd=1, reward_fn=reward)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    smp_job.eps_transform = None
    tt = torch.tensor([times[-1]], device=dev)
    same = torch.tensor([float(xh[:, :3].double().sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sums = [torch.zeros_like(same) for _ in range(world)]
        dist.all_gather(sums, same)
        identical = all(float(x) == float(sums[0]) for x in sums)
    else:
        identical = True
    job['atp'] = {'ligands': Ba, 'candidate_groups': 5, 'events': 6, 'ranks': world, 'seconds': float(tt), 'value': Ba / float(tt),
                  'unit': 'ligands/s', 'first_run_seconds': times[0], 'ranks_end_identical': bool(identical),
                  'collective': 'one all_gather per event (scores + pocket positions + latents of the rank\'s groups)' if world > 1 else None}
    return job


def main():
    global POCKET_ATOMS
    args = parse()
    if args.pocket_atoms:
        POCKET_ATOMS = args.pocket_atoms
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
