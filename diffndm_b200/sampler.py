"""Host-side mirror of the reference sampler surface around the CUDA denoiser.

Mirrors (names, argument meaning, outputs) the parts of ``ConditionalDDPM`` (reference
equivariant_diffusion/conditional_model.py) that sit on the hot path:

* ``sample_p_zs_given_zt``      (:483-540)   one reverse-diffusion step
* ``sample_p_xh_given_z0``      (:136-160)   final p(x, h | z_0) head
* ``my_to_x0``                  (:457-468)   x0 look-ahead used by SPSA / ATP
* ``my_update_z_lig``           (:760-813)   SPSA guidance, with the 2k perturbed copies batched into two
                                             denoiser calls instead of 4k sequential ones
* ``sample_given_pocket``       (:886-1489)  the sampling loop (plain / SPSA / ATP)

Loop control, schedule scalars and reward bookkeeping stay Python; every tensor operation is a call into the
C-ABI engine (``dndm_egnn_forward`` + ``dndm_sampler_step``).  Host chemistry (RDKit / OpenBabel scoring)
stays outside: guidance takes a ``reward_fn(x_lig, atom_types, lig_mask) -> list[float]`` callable, exactly the
role ``handle_to_mol`` + ``my_reward_for_SPSA`` / ``my_reward_for_SVDD`` play in the reference.

The reference's per-step host syncs (``.item()`` in assert_mean_zero_with_mask, ``empty_cache()``) are replaced
by sticky device flags checked once per trajectory (or per step with ``check_every_step=True``).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from .engine import B200EGNNDynamics, FLAG_COM_DRIFT, FLAG_EDGE_OVERFLOW, FLAG_NAN


def _clip_noise_schedule(alphas2, clip_value=0.001):
    alphas2 = np.concatenate([np.ones(1), alphas2], axis=0)
    alphas_step = np.clip(alphas2[1:] / alphas2[:-1], a_min=clip_value, a_max=1.)
    return np.cumprod(alphas_step, axis=0)


def polynomial_gamma(timesteps: int, precision: float, power: float) -> torch.Tensor:
    """gamma[T+1] of PredefinedNoiseSchedule('polynomial_<power>') -- en_diffusion.py:1146-1191."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas2 = (1 - np.power(x / steps, power)) ** 2
    alphas2 = _clip_noise_schedule(alphas2, clip_value=0.001)
    alphas2 = (1 - 2 * precision) * alphas2 + precision
    sigmas2 = 1 - alphas2
    return torch.from_numpy(-(np.log(alphas2) - np.log(sigmas2))).float()


def segment_sum_sorted(x: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    """Per-segment sums of rows whose segment ids are sorted (batch masks, utils.py:145-153) WITHOUT atomics: an fp64 prefix
    sum differenced at the segment boundaries.  ``index_add_`` on CUDA adds in whatever order its atomics land, which made
    two trajectories from the same seed differ in the last bits (amplified by the reverse process); the per-sample means the
    sampler takes (pocket centre, perturbation centring, inpainting anchors) therefore go through this."""
    cnt = torch.bincount(idx, minlength=n)
    ends = torch.cumsum(cnt, 0)
    cs = torch.cumsum(x.double(), dim=0)
    cs = torch.cat([torch.zeros_like(cs[:1]), cs], dim=0)
    return (cs[ends] - cs[ends - cnt]).to(x.dtype)


def segment_mean_sorted(x: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    cnt = torch.bincount(idx, minlength=n).clamp(min=1).to(x.dtype)
    return segment_sum_sorted(x, idx, n) / cnt[:, None]


def _event_fn(reward_fn, kind: str):
    """The reference scores SPSA rounds with my_reward_for_SPSA and ATP selections with my_reward_for_SVDD
    (conditional_model.py:743-744, 1187, 1200): a reward object may offer both through ``for_event(kind)``
    (rewards_rdkit.GuidanceRewards); a plain callable serves every event."""
    return reward_fn.for_event(kind) if hasattr(reward_fn, 'for_event') else reward_fn


class NoiseProvider:
    """Source of the Gaussian draws of a guided trajectory, for parity tests that replay the reference's draws.  The
    sampler batches what the reference does sequentially, so it asks for the draws of a whole event at once; an
    implementation hands them out in whatever order its source recorded them (tests/golden/guidance_common.py walks the
    reference's torch.randn sequence).  All returns are float32 arrays / tensors, D = 3 + atom_nf.

    step(n_l)                  -> [n_l, D]                       z_T, one reverse step, the final p(x, h | z_0) head
    spsa(k, sizes)             -> ([k, n_l, 3], [2k, n_l, D])    raw perturbation draws (one per molecule and round in the
                                                                 reference, :771-782) and the x0 look-ahead draws, +U rounds
                                                                 first, then -U rounds
    atp(n_groups, n_l)         -> ([G-1, n_l, D], [G, n_l, D])   extra-candidate reverse steps; x0 look-aheads (group 0 first)
    mixed(n_groups, k, sizes)  -> ([G-1, n_l, D], [(spsa draws)] * (G-1), [G, n_l, D])   the s == 30 branch: per extra candidate
                                                                 its reverse step, its SPSA draws, its x0 look-ahead
    """

    def step(self, n_l):
        raise NotImplementedError

    def spsa(self, k, sizes):
        raise NotImplementedError

    def atp(self, n_groups, n_l):
        raise NotImplementedError

    def mixed(self, n_groups, k, sizes):
        raise NotImplementedError


class _GraphedReverseStep:
    """One unguided reverse step (denoiser forward + noise draw + fused p(z_s|z_t) update, in place on the state buffers)
    captured once in a CUDA graph and replayed with new (t, coefficient) values -- removes the ~60 kernel launches and the
    per-call host work from every step of a trajectory.  Valid while the mask tensors and the batch layout stay the same
    (between ATP events); flags are sticky on the device and read by the caller."""

    def __init__(self, sampler: 'ConditionalSampler', z_lig, xh_pocket, lig_mask, pocket_mask, B):
        eng = sampler.engine
        dev = sampler.device
        self.engine = eng
        self.z, self.xp = z_lig, xh_pocket                     # updated in place
        self.lig_mask, self.pocket_mask = lig_mask, pocket_mask   # the graph reads these buffers on every replay
        # Per-step scalars without per-step host work: row s of `table` holds (t, step coefficients) of step s already expanded
        # over the samples; the captured step gathers row `s_idx` into `packed` -- t_buf / coef_buf are views of it -- and
        # decrements `s_idx`, so a run of consecutive steps is one graph replay per step and nothing else (two expand-copies
        # per step before: 28 us of a 1.65 ms step).  The host rewrites `s_idx` only when a step is not the successor of the
        # previous replay (first step, after a guidance event).
        Bp = (B + 3) // 4 * 4                                   # keeps coef_buf 16-byte aligned
        self.table = torch.zeros((sampler.T, Bp + 3 * B), device=dev)
        self.packed = torch.zeros((1, Bp + 3 * B), device=dev)
        self.t_buf = self.packed[0, :B].view(B, 1)
        self.coef_buf = self.packed[0, Bp:].view(B, 3)
        self.s_idx = torch.zeros(1, dtype=torch.long, device=dev)
        self._next_s = None
        self._tables_of = None
        # the captured kernels read the transform's tensors on every replay: the graph is valid for THIS transform object only
        # (part of the cache key) and keeps it alive
        self.transform = transform = sampler.eps_transform

        def body():
            torch.index_select(self.table, 0, self.s_idx, out=self.packed)
            self.s_idx.sub_(1)
            eps, _ = eng.forward(self.z, self.xp, self.t_buf, lig_mask, pocket_mask, B, want_pocket=False)
            if transform is not None:
                eps = transform(eps, self.z, self.xp, self.t_buf, lig_mask, pocket_mask)
            nz = torch.randn_like(self.z)
            eng.sampler_step(self.z, eps, nz, self.xp, self.coef_buf, lig_mask, pocket_mask, B, z_out=self.z,
                             pocket_out=self.xp, check_com=True)

        keep_z, keep_p = self.z.clone(), self.xp.clone()
        rng = torch.cuda.get_rng_state(dev)                    # the dry runs below must not consume the caller's noise stream
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()                                             # warm-up outside capture (allocator, layout cache); row 0 of the
            self.s_idx.zero_()                                 # still empty table: all-zero scalars, state restored below
        torch.cuda.current_stream().wait_stream(side)
        # the captured step derives the per-sample offsets from the masks itself (cache reset -> the first prepare_batch of
        # the body launches the two mask kernels inside the graph): a replay does not depend on what other calls left in
        # the engine's scratch buffers
        eng.set_static_masks(True)
        # Capture through capture_begin / capture_end on the side stream instead of the `torch.cuda.graph` context manager: its
        # __enter__ runs gc.collect() + torch.cuda.empty_cache(), and a job captures one graph per pocket shape while older
        # graphs are evicted -- the cudaFree of their pools made a capture take 0.2 - 1.9 s now and then (10 ms otherwise).
        # All graphs of a sampler share one private pool (they never replay concurrently and keep nothing in it between
        # replays: the state buffers live outside), so an evicted graph's memory is reused, not returned to the driver.
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            self.graph.capture_begin(pool=sampler.graph_pool)
            try:
                body()
            finally:
                self.graph.capture_end()
        torch.cuda.current_stream().wait_stream(side)
        self.z.copy_(keep_z)
        self.xp.copy_(keep_p)
        torch.cuda.synchronize(dev)
        torch.cuda.set_rng_state(rng, dev)
        eng.read_flags(consume=FLAG_COM_DRIFT)                 # the dry runs may have tripped the COM-drift flag (state restored
                                                               # above); NaN / overflow bits of earlier calls stay pending

    def set_tables(self, t_all: torch.Tensor, coef_all: torch.Tensor):
        """Per-step scalars of a trajectory: t_all [T'], coef_all [T', 3] (device, T' <= sampler.T; the same for every sample of
        an unguided step).  Once per trajectory."""
        if self._tables_of is t_all:
            return
        n, B = int(t_all.shape[0]), self.t_buf.shape[0]
        assert n <= self.table.shape[0], 'more steps than the sampler was built for'
        self.table[:n, :B] = t_all[:, None]
        self.table[:n, self.packed.shape[1] - 3 * B:] = coef_all[:, None, :].expand(n, B, 3).reshape(n, 3 * B)
        self._tables_of = t_all
        self._next_s = None

    def __call__(self, s: int):
        """The reverse step s + 1 -> s with the scalars of table row s."""
        if self._next_s != s:
            self.s_idx.fill_(s)
        self.graph.replay()
        self._next_s = s - 1
        self.engine.set_static_masks(True)      # host-side cache reset: the device scratch now holds THIS graph's layout


class ConditionalSampler:
    """Sampler over a ``B200EGNNDynamics``; all state lives on the GPU."""

    def __init__(self, dynamics: B200EGNNDynamics, timesteps: int = 500, noise_schedule: str = 'polynomial_2',
                 noise_precision: float = 5.0e-4, norm_values=(1.0, 4.0), norm_biases=(None, 0.0),
                 check_every_step: bool = False, gamma_table: Optional[torch.Tensor] = None):
        assert not dynamics.update_pocket_coords           # conditional_model.py:24
        self.dynamics = dynamics
        self.engine = dynamics.engine
        self.T = timesteps
        self.n_dims = 3
        self.atom_nf = dynamics.cfg.atom_nf
        self.residue_nf = dynamics.cfg.residue_nf      # 20 for C-alpha pockets (amino-acid one-hot), else = atom_nf
        self.norm_values = norm_values
        self.norm_biases = norm_biases
        self.check_every_step = check_every_step
        self.overlap_scoring = True                # SPSA: score one half of a round on the host while the GPU denoises the other
        self._graph_cache = {}                     # (weights version, B, N_l, N_p, transform) -> _GraphedReverseStep
        self.graph_pool = torch.cuda.graph_pool_handle()   # one private memory pool for every captured reverse step
        # optional device-side hook applied to every denoiser output: eps = f(eps, z, xh_pocket, t [B,1] on the device,
        # lig_mask, pocket_mask).  Capture-safe callables only (it runs inside the graphed reverse step as well).  Used by
        # the benchmarks for the synthetic score that stands in for trained weights (synthetic.PointMassScore).
        self.eps_transform = None
        self.atp_group = None                      # torch.distributed group over which ATP candidate groups are split
        self.cand_gen = None                       # generator of the candidate / perturbation draws (None: default generator)
        if gamma_table is not None:               # a live reference schedule: PredefinedNoiseSchedule.gamma, en_diffusion.py:1189-1191
            self.gamma = torch.as_tensor(gamma_table).detach().cpu().float().reshape(-1)
            assert self.gamma.numel() == timesteps + 1
        else:
            assert noise_schedule.startswith('polynomial_')
            self.gamma = polynomial_gamma(timesteps, noise_precision, float(noise_schedule.split('_')[1]))   # CPU fp32
        self.device = torch.device('cuda', self.engine.device)
        self._build_tables()

    # -- schedule scalars (en_diffusion.py:83-108, 870-883), same torch fp32 ops as the reference, once ----------
    def _build_tables(self):
        g = self.gamma
        sig = lambda gm: torch.sqrt(torch.sigmoid(gm))
        alp = lambda gm: torch.sqrt(torch.sigmoid(-gm))
        self.sigma_tab = sig(g)
        self.alpha_tab = alp(g)
        self.gamma_dev = g.to(self.device)

    def _h2d(self, t: torch.Tensor) -> torch.Tensor:
        """Host -> device without blocking the host: a blocking ``.to(device)`` synchronises the stream, i.e. waits for every
        kernel queued so far, which serialises the host with the GPU (and with it the overlapped host scoring).  The pinned
        staging block is recycled by torch's host allocator only after the copy has run."""
        if t.is_cuda:
            return t.to(self.device)
        return t.contiguous().pin_memory().to(self.device, non_blocking=True)

    def _const_rows(self, row, n: int) -> torch.Tensor:
        """[n, len(row)] device tensor with every row = ``row``, built by fill kernels (no host copy, no sync)."""
        out = torch.empty((n, len(row)), device=self.device)
        for j, v in enumerate(row):
            out[:, j] = float(v)
        return out

    def lookup(self, t: torch.Tensor) -> torch.Tensor:
        """gamma(t) with t in [0,1] -- PredefinedNoiseSchedule.forward, en_diffusion.py:1193-1195 (CPU table)."""
        return self.gamma[torch.round(t.detach().cpu().float() * self.T).long().reshape(-1)]

    def step_coefficients(self, gamma_s: torch.Tensor, gamma_t: torch.Tensor) -> torch.Tensor:
        """[B,3] = (1/alpha_ts, sigma2_ts/alpha_ts/sigma_t, sigma_ts*sigma_s/sigma_t) -- conditional_model.py:486-529."""
        sigma2_ts = -torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t))
        alpha_ts = torch.exp(0.5 * (F.logsigmoid(-gamma_t) - F.logsigmoid(-gamma_s)))
        sigma_ts = torch.sqrt(sigma2_ts)
        sigma_s = torch.sqrt(torch.sigmoid(gamma_s))
        sigma_t = torch.sqrt(torch.sigmoid(gamma_t))
        return torch.stack([1.0 / alpha_ts, sigma2_ts / alpha_ts / sigma_t, sigma_ts * sigma_s / sigma_t], dim=1)

    # -- elementary moves -------------------------------------------------------------------------------------------
    def _eps(self, z_lig, xh_pocket, t, lig_mask, pocket_mask, B):
        """Ligand part of the denoiser output.  Every conditional call site discards the pocket part (`eps, _ = ...`,
        conditional_model.py:143, 458, 504), so it is not computed: the last block then skips the pocket atoms that no
        ligand atom reads (bit-identical ligand output)."""
        keep = self.dynamics.compute_pocket_output
        self.dynamics.compute_pocket_output = False
        try:
            eps = self.dynamics(z_lig, xh_pocket, t, lig_mask, pocket_mask, n_samples=B)[0]
        finally:
            self.dynamics.compute_pocket_output = keep
        if self.eps_transform is not None:
            eps = self.eps_transform(eps, z_lig, xh_pocket, t.to(self.device).reshape(-1, 1), lig_mask, pocket_mask)
        return eps

    def _noise(self, n, noise=None):
        if noise is not None:
            return noise.to(self.device, torch.float32)
        return torch.randn((n, self.n_dims + self.atom_nf), device=self.device)    # sample_gaussian, en_diffusion.py:957-960

    def sample_p_zs_given_zt(self, s, t, zt_lig, xh0_pocket, ligand_mask, pocket_mask, noise=None, n_samples=None):
        """conditional_model.py:483-540 (optimize=0; AdjustNet is training-only and off the sampling path)."""
        B = int(n_samples if n_samples is not None else t.numel())
        coef = self._h2d(self.step_coefficients(self.lookup(s), self.lookup(t)))
        eps = self._eps(zt_lig, xh0_pocket, self._h2d(t), ligand_mask, pocket_mask, B)
        zs, xp = self.engine.sampler_step(zt_lig, eps, self._noise(len(ligand_mask), noise), xh0_pocket, coef,
                                          ligand_mask, pocket_mask, B, check_com=True)     # assert on z_t, :535
        if self.check_every_step:
            self._raise_on_flags()
        return zs, xp

    def sample_p_xh_given_z0(self, z0_lig, xh0_pocket, lig_mask, pocket_mask, batch_size, noise=None):
        """conditional_model.py:136-160.  Returns x_lig, one-hot h_lig (int64), x_pocket, h_pocket."""
        B = int(batch_size)
        g0 = self.lookup(torch.zeros((B, 1)))                           # CPU table: no device round trip, no host sync
        t0 = torch.zeros((B, 1), device=self.device)
        eps0 = self._eps(z0_lig, xh0_pocket, t0, lig_mask, pocket_mask, B)
        sigma_x = torch.exp(0.5 * g0)                                   # SNR(-0.5 gamma_0)
        sigma0, alpha0 = torch.sqrt(torch.sigmoid(g0)), torch.sqrt(torch.sigmoid(-g0))
        coef = self._h2d(torch.stack([1.0 / alpha0, sigma0 / alpha0, sigma_x], dim=1))        # compute_x_pred
        xh, xp = self.engine.sampler_step(z0_lig, eps0, self._noise(len(lig_mask), noise), xh0_pocket, coef, lig_mask,
                                          pocket_mask, B)
        x_lig = xh[:, :3] * self.norm_values[0]
        h_lig = z0_lig[:, 3:] * self.norm_values[1] + self.norm_biases[1]
        x_pocket = xp[:, :3] * self.norm_values[0]
        h_pocket = xp[:, 3:] * self.norm_values[1] + self.norm_biases[1]
        h_lig = F.one_hot(torch.argmax(h_lig, dim=1), self.atom_nf)
        return x_lig, h_lig, x_pocket, h_pocket

    def my_to_x0(self, t, zt_lig, xh0_pocket, ligand_mask, pocket_mask, n_samples, noise=None):
        """conditional_model.py:457-468: x0 look-ahead (two denoiser calls)."""
        B = int(n_samples)
        eps_t = self._eps(zt_lig, xh0_pocket, self._h2d(t), ligand_mask, pocket_mask, B)
        gt = self.lookup(t)
        alpha_t = self._h2d(torch.exp(0.5 * F.logsigmoid(-gt)))
        sigma_t = self._h2d(torch.sqrt(torch.sigmoid(gt)))
        z0 = (zt_lig - sigma_t[ligand_mask][:, None] * eps_t) / alpha_t[ligand_mask][:, None]
        return self.sample_p_xh_given_z0(z0, xh0_pocket, ligand_mask, pocket_mask, B, noise=noise)

    def remove_mean_batch(self, x_lig, x_pocket, lig_mask, pocket_mask, n_samples):
        """conditional_model.py:1793-1801 through the fused kernel (coef = identity)."""
        B = int(n_samples)
        z = torch.zeros((x_lig.shape[0], 3 + self.atom_nf), device=self.device)
        z[:, :3] = x_lig
        p = torch.zeros((x_pocket.shape[0], 3 + self.residue_nf), device=self.device)
        p[:, :3] = x_pocket
        coef = self._const_rows((1.0, 0.0, 0.0), B)
        zo, po = self.engine.sampler_step(z, None, z, p, coef, lig_mask, pocket_mask, B)
        return zo[:, :3], po[:, :3]

    def set_distributed_atp(self, group, shared_seed: int, rank_seed: Optional[int] = None):
        """Split the ATP candidate groups of ONE trajectory over the ranks of ``group``.  The ranks must carry the same
        trajectory state, so everything on the common path -- z_T, the reverse steps, the SPSA perturbations, the s == 30
        chain -- draws from the default CUDA generator, seeded identically here on every rank, while the extra ATP
        candidates of a rank come from its own generator (seed ``rank_seed``, default shared_seed + 1 + rank): identical
        seeds there would make every rank draw the same candidates.  The engine is deterministic, so equal inputs and
        equal draws keep the states bit-identical between events."""
        self.atp_group = group
        torch.cuda.manual_seed(int(shared_seed))
        torch.manual_seed(int(shared_seed))
        self.cand_gen = torch.Generator(device=self.device)
        r = dist.get_rank(group) if group is not None else 0
        self.cand_gen.manual_seed(int(shared_seed) + 1 + r if rank_seed is None else int(rank_seed))

    def _raise_on_flags(self):
        flags = self.engine.read_flags()
        if flags & FLAG_EDGE_OVERFLOW:
            raise RuntimeError('diffndm_b200: edge capacity exceeded (raise max_edges)')
        if flags & FLAG_NAN:
            raise ValueError("NaN detected in EGNN output")                       # dynamics.py:155-159
        if flags & FLAG_COM_DRIFT:
            raise AssertionError('Mean is not zero')                              # en_diffusion.py:930-935

    # -- SPSA (conditional_model.py:724-813), 2k perturbed copies batched -------------------------------------------
    def my_update_z_lig(self, z_lig, xh_pocket, lig_mask, pocket_mask, t_array, n_samples, zeta, reward_fn,
                        guidance_scale=1e-3, k=10, perturbations=None, x0_noise=None, perturbation_noise=None,
                        generator=None):
        """Symmetric finite-difference guidance.  The reference runs 2k x my_to_x0 sequentially (:764-800); here the
        2k copies are concatenated along the batch axis: ONE denoiser call at t and ONE at t=0 on 2k*B samples.
        For parity tests ``perturbation_noise`` [k, N_l, 3] (the raw draws, centred and scaled here) or ``perturbations``
        [k, N_l, 3] (already zeta * centred) and ``x0_noise`` [2k, N_l, 13] (+U rounds first) may be injected."""
        B, n_l, n_p = int(n_samples), z_lig.shape[0], xh_pocket.shape[0]
        reward_fn = _event_fn(reward_fn, 'spsa')
        sizes = torch.bincount(lig_mask, minlength=B)
        if perturbations is None:                                                   # my_perturbation_for_molecule :724-736
            noise = (torch.randn((k, n_l, 3), device=self.device, generator=generator) if perturbation_noise is None
                     else self._h2d(torch.as_tensor(perturbation_noise, dtype=torch.float32)))
            mean = segment_sum_sorted(noise.permute(1, 0, 2).reshape(n_l, k * 3), lig_mask, B).reshape(B, k, 3).permute(1, 0, 2) \
                / sizes[None, :, None]
            perturbations = zeta * (noise - mean[:, lig_mask])
        U = self._h2d(perturbations)
        reps = 2 * k
        z_rep = z_lig.unsqueeze(0).repeat(reps, 1, 1)
        z_rep[:k, :, :3] += U
        z_rep[k:, :, :3] -= U
        offs = (torch.arange(reps, device=self.device) * B)
        big_lig_mask = (lig_mask.unsqueeze(0) + offs[:, None]).reshape(-1)
        big_pocket_mask = (pocket_mask.unsqueeze(0) + offs[:, None]).reshape(-1)
        big_pocket = xh_pocket.unsqueeze(0).repeat(reps, 1, 1).reshape(reps * n_p, -1)
        # kept on the host: my_to_x0 looks the schedule up in the CPU table, and a device tensor there would cost a
        # device->host copy that blocks the host until everything queued so far has run
        big_t = t_array.detach().cpu().reshape(1, B, 1).repeat(reps, 1, 1).reshape(reps * B, 1)
        if x0_noise is None and generator is not None:
            x0_noise = torch.randn((reps, n_l, self.n_dims + self.atom_nf), device=self.device, generator=generator)
        nz = None if x0_noise is None else self._h2d(torch.as_tensor(x0_noise, dtype=torch.float32)).reshape(reps * n_l, -1)
        if hasattr(reward_fn, 'submit') and self.overlap_scoring:
            # host scoring overlapped with GPU denoising: the +U copies are denoised first and go to the scorer's worker
            # processes (device->host copy on its side stream, gated by an event) while the GPU denoises the -U copies
            z_flat = z_rep.reshape(reps * n_l, -1)
            pending = []
            for half in range(2):
                a_l, b_l = half * k * n_l, (half + 1) * k * n_l
                a_p, b_p = half * k * n_p, (half + 1) * k * n_p
                a_b, b_b = half * k * B, (half + 1) * k * B
                x_h, h_h, _, _ = self.my_to_x0(big_t[a_b:b_b], z_flat[a_l:b_l], big_pocket[a_p:b_p],
                                               big_lig_mask[a_l:b_l] - a_b, big_pocket_mask[a_p:b_p] - a_b, k * B,
                                               noise=None if nz is None else nz[a_l:b_l])
                types_h, mask_h = h_h.argmax(1), big_lig_mask[a_l:b_l] - a_b
                done = torch.cuda.Event()          # recorded after EVERYTHING the scorer's copy stream will read
                done.record()
                pending.append((x_h, types_h, mask_h, done))
                if half == 1:                      # both halves are queued: hand them over in order
                    handles = [reward_fn.submit(x, t, m, after=ev) for x, t, m, ev in pending]
            rewards = torch.as_tensor(handles[0].result() + handles[1].result(), dtype=torch.float32,
                                      device=self.device).reshape(reps, B)
        else:
            x_l, h_l, _, _ = self.my_to_x0(big_t, z_rep.reshape(reps * n_l, -1), big_pocket, big_lig_mask, big_pocket_mask,
                                           reps * B, noise=nz)
            rewards = torch.as_tensor(reward_fn(x_l, h_l.argmax(1), big_lig_mask), dtype=torch.float32,
                                      device=self.device).reshape(reps, B)
        f_plus, f_minus = rewards[:k], rewards[k:]
        self.last_spsa_rewards = (f_plus, f_minus)                                  # introspection (tests, logging)
        dd = (f_plus - f_minus) / (2 * 1e-4)                                         # hard-coded divisor, :799
        grad = (dd[:, lig_mask, None] * U).sum(0) / k                                # :749-758, :801 (sum, then / len)
        # x += guidance_scale * grad ; COM removal for ligand and pocket  (:803-812)
        coef = self._const_rows((1.0, 0.0, 0.0), B)
        return self.engine.sampler_step(z_lig, None, z_lig, xh_pocket, coef, lig_mask, pocket_mask, B, grad=grad,
                                        lam=float(guidance_scale))

    def _unnormalize_quirk(self, z_lig, xh_pocket, lig_mask, pocket_mask, B):
        """The reference rescales features by norm_values[1] after every SPSA / ATP event
        (conditional_model.py:1235-1240, 1253-1258) -- reproduced, not fixed."""
        z = z_lig.clone()
        p = xh_pocket.clone()
        z[:, 3:] = z[:, 3:] * self.norm_values[1] + self.norm_biases[1]
        p[:, 3:] = p[:, 3:] * self.norm_values[1] + self.norm_biases[1]
        z[:, :3], p[:, :3] = self.remove_mean_batch(z[:, :3] * self.norm_values[0], p[:, :3] * self.norm_values[0],
                                                    lig_mask, pocket_mask, B)
        return z, p

    # -- the sampling loop -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample_given_pocket(self, pocket, num_nodes_lig, timesteps: Optional[int] = None, svdd: int = 0, spsa: int = 0,
                            reward_fn: Optional[Callable] = None, noise=None,
                            spsa_schedule=(30, 2), svdd_schedule=(50, 10), svdd_groups: int = 5, spsa_k: int = 10,
                            use_cuda_graph: bool = True, mixed_at: Optional[int] = 30, resume=None, stop_after: int = 0):
        """ConditionalDDPM.sample_given_pocket (conditional_model.py:886-1489) without the host chemistry arguments.

        ``pocket``: dict with 'x' [N_p,3], 'one_hot' [N_p,residue_nf], 'size' [B], 'mask' [N_p] (prepare_pocket layout,
        lightning_modules.py:763-801).  ``noise``: either a tensor [timesteps+2, N_l, 13] with the Gaussian draws of an
        UNGUIDED run in the reference's order (z_T, one per step, final head) or a ``NoiseProvider`` (guided runs).
        Event order inside one step follows the reference: reverse step, ATP event (svdd, s <= 50, s % 10 == 0; :1085-1241),
        SPSA update (spsa, s <= 30, s % 2 == 0; :1243-1259) and, at s == ``mixed_at`` (30) with spsa on, the mixed branch
        (:1261-1418); every event is followed by the reference's feature rescaling (:1235-1240).
        ``use_cuda_graph``: replay the unguided reverse step from a CUDA graph (re-captured after every ATP event, whose
        re-batching changes the masks); the NaN / COM-drift flags are then checked at guidance events and at the end
        instead of at the failing step (``check_every_step=True`` keeps the reference's per-step behaviour, eagerly).
        ``resume`` = (z_lig, xh_pocket, lig_mask, s_next): continue a trajectory from a saved state (checkpoint / resume);
        ``stop_after`` = s: return the latent state (z_lig, xh_pocket, lig_mask, pocket_mask) after the events of step s.
        Returns (xh_lig [N_l, 3+atom_nf] with one-hot features, xh_pocket, lig_mask, pocket_mask) like the reference.
        """
        timesteps = self.T if timesteps is None else timesteps
        dev = self.device
        B = len(pocket['size'])
        provider = noise if isinstance(noise, NoiseProvider) else None
        table = None if provider is not None else noise
        x_p = pocket['x'].to(dev, torch.float32) / self.norm_values[0]
        h_p = (pocket['one_hot'].to(dev).float() - self.norm_biases[1]) / self.norm_values[1]
        pocket_mask = pocket['mask'].to(dev).long()
        xh0_pocket = torch.cat([x_p, h_p], dim=1).contiguous()
        step = 0

        def step_noise(n):
            if provider is not None:
                return self._h2d(torch.as_tensor(provider.step(n), dtype=torch.float32))
            return None if table is None else table[step]

        if resume is None:
            sizes = torch.as_tensor(num_nodes_lig, device=dev).long()
            lig_mask = torch.repeat_interleave(torch.arange(B, device=dev), sizes)          # utils.py:145-153
            n_l = int(lig_mask.numel())
            # z_T ~ N(pocket COM, I), projected to the ligand-COM-free subspace (:923-930)
            mu_x = segment_mean_sorted(x_p, pocket_mask, B)
            mu = torch.cat([mu_x, torch.zeros((B, self.atom_nf), device=dev)], dim=1)[lig_mask].contiguous()
            ident = torch.tensor([[1.0, 0.0, 1.0]], device=dev).repeat(B, 1)
            z_lig, xh_pocket = self.engine.sampler_step(mu, None, self._noise(n_l, step_noise(n_l)), xh0_pocket, ident,
                                                        lig_mask, pocket_mask, B)
            s_first = timesteps - 1
        else:
            z_lig, xh_pocket, lig_mask, s_first = resume
            z_lig = z_lig.to(dev, torch.float32).contiguous()
            xh_pocket = xh_pocket.to(dev, torch.float32).contiguous()
            lig_mask = lig_mask.to(dev).long()
            step = timesteps - 1 - int(s_first)
        self.engine.set_static_masks(True)          # lig_mask / pocket_mask are fixed tensors between ATP events
        graphed = use_cuda_graph and noise is None and not self.check_every_step
        gstep = None
        if graphed:                                 # per-step scalars of the whole trajectory, on the device, once
            s_all = torch.arange(timesteps, dtype=torch.float32)
            coef_all = self.step_coefficients(self.lookup(s_all / timesteps), self.lookup((s_all + 1) / timesteps)).to(dev)
            t_all = ((s_all + 1) / timesteps).to(dev)
        nan_check, self.dynamics.check_nan = self.dynamics.check_nan, (self.dynamics.check_nan and not graphed)
        try:
            for s in reversed(range(stop_after, int(s_first) + 1)):
                s_array = torch.full((B, 1), fill_value=s, dtype=torch.float32) / timesteps
                t_array = torch.full((B, 1), fill_value=s + 1, dtype=torch.float32) / timesteps
                step += 1
                # inside an ATP window the batch is re-built every few steps (other ligand sizes, another mask): capturing a graph
                # per event costs more than the launches it saves, so those steps run eagerly (without per-step host syncs).
                # An SPSA update keeps shapes and masks: the cached graph is reused, the new state is copied into its buffers.
                in_window = svdd == 1 and s <= svdd_schedule[0]
                if graphed and not in_window:
                    if gstep is None:
                        # graphs are kept across trajectories: the same pocket with the same ligand sizes (the usual
                        # "n_samples per pocket" loop) replays the graph captured for the first batch
                        key = (self.engine.weights_version, B, int(z_lig.shape[0]), int(xh_pocket.shape[0]), id(self.eps_transform))
                        gstep = self._graph_cache.get(key)
                        if gstep is not None and gstep.transform is self.eps_transform and torch.equal(gstep.lig_mask, lig_mask) \
                                and torch.equal(gstep.pocket_mask, pocket_mask):
                            gstep.z.copy_(z_lig)
                            gstep.xp.copy_(xh_pocket)
                        else:
                            gstep = _GraphedReverseStep(self, z_lig.contiguous().clone(), xh_pocket.contiguous().clone(),
                                                        lig_mask, pocket_mask, B)
                            if len(self._graph_cache) >= 4:
                                self._graph_cache.pop(next(iter(self._graph_cache)))
                            self._graph_cache[key] = gstep
                        z_lig, xh_pocket = gstep.z, gstep.xp
                        gstep.set_tables(t_all, coef_all)
                    gstep(s)
                else:
                    z_lig, xh_pocket = self.sample_p_zs_given_zt(s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask,
                                                                 noise=step_noise(int(lig_mask.numel())), n_samples=B)
                if svdd == 1 and s <= svdd_schedule[0] and s % svdd_schedule[1] == 0:
                    if graphed:
                        self._raise_on_flags()
                    self.engine.set_static_masks(False)     # candidate batches use other masks; the winners get a new one
                    z_lig, xh_pocket, lig_mask = self._atp_event(s, s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask,
                                                                 B, reward_fn, svdd_groups, provider=provider)
                    z_lig, xh_pocket = self._unnormalize_quirk(z_lig, xh_pocket, lig_mask, pocket_mask, B)
                    self.engine.set_static_masks(True)
                    gstep = None                            # new state tensors and ligand mask: capture again
                if spsa == 1 and s <= spsa_schedule[0] and s % spsa_schedule[1] == 0:
                    if graphed:
                        self._raise_on_flags()
                    zeta = 1e-3 * (s / 500)                                                   # :1244-1245
                    pn, xn = self._spsa_draws(provider, spsa_k, lig_mask, B)
                    z_lig, xh_pocket = self.my_update_z_lig(z_lig, xh_pocket, lig_mask, pocket_mask, t_array, B, zeta,
                                                            reward_fn, guidance_scale=1e-3, k=spsa_k, perturbation_noise=pn,
                                                            x0_noise=xn)
                    z_lig, xh_pocket = self._unnormalize_quirk(z_lig, xh_pocket, lig_mask, pocket_mask, B)
                    if mixed_at is not None and s == mixed_at:                                # :1261-1418
                        self.engine.set_static_masks(False)
                        z_lig, xh_pocket, lig_mask = self._mixed_event(s, s_array, t_array, z_lig, xh_pocket, lig_mask,
                                                                       pocket_mask, B, reward_fn, svdd_groups, zeta, 1e-3,
                                                                       spsa_k, provider=provider)
                        z_lig, xh_pocket = self._unnormalize_quirk(z_lig, xh_pocket, lig_mask, pocket_mask, B)
                        self.engine.set_static_masks(True)
                    gstep = None
            if stop_after > 0:
                self._raise_on_flags()
                return z_lig, xh_pocket, lig_mask, pocket_mask
            step += 1
            x_lig, h_lig, x_pocket, h_pocket = self.sample_p_xh_given_z0(
                z_lig, xh_pocket, lig_mask, pocket_mask, B, noise=step_noise(int(lig_mask.numel())))
        finally:
            self.dynamics.check_nan = nan_check
            self.engine.set_static_masks(False)
        self._raise_on_flags()
        # CoG drift correction (:1431-1438)
        cog = segment_sum_sorted(x_lig, lig_mask, B).abs().max().item()
        if cog > 5e-2:
            x_lig, x_pocket = self.remove_mean_batch(x_lig, x_pocket, lig_mask, pocket_mask, B)
        return torch.cat([x_lig, h_lig.float()], dim=1), torch.cat([x_pocket, h_pocket], dim=1), lig_mask, pocket_mask

    def _spsa_draws(self, provider, k, lig_mask, B):
        if provider is None:
            return None, None
        sizes = torch.bincount(lig_mask, minlength=B).tolist()
        return provider.spsa(k, sizes)

    # -- inpainting (RePaint resampling), conditional_model.py:1491-1790 ----------------------------------------------------
    @torch.no_grad()
    def inpaint(self, ligand, pocket, lig_fixed, svdd: int = 0, resamplings: int = 1, timesteps: Optional[int] = None,
                center: str = 'ligand', reward_fn: Optional[Callable] = None, noise=None, spsa_window=(12, 16),
                spsa_k: int = 10, svdd_schedule=(10, 2), svdd_groups: int = 5):
        """ConditionalDDPM.inpaint without the host-chemistry arguments: keep the atoms flagged in ``lig_fixed`` [N_l],
        generate the rest.  ``ligand``: dict with 'x' [N_l,3], 'one_hot' [N_l,atom_nf], 'size' [B], 'mask' [N_l];
        ``pocket`` as in sample_given_pocket.  Every (s, u) iteration is: reverse step (:1566-1568), forward-noised known
        part q(z_s | x) on the pocket-shifted input (:1588-1594), COM matching over the fixed atoms + blend (:1596-1609)
        and, except on the last resampling, the re-noising move z_s -> z_t (:1611-1615).  All three stochastic moves run
        in the fused sampler-step kernel (out = c0 z - c1 eps + c2 noise, then the ligand-COM projection).
        The reference hard-wires an SPSA update for 12 <= s <= 16 on the first resampling (:1570-1586) and an ATP event at
        s <= 10, s % 2 == 0 when svdd == 1 (:1627-1778); both need host rewards and run only if ``reward_fn`` is given.
        ``noise``: optional iterable of [N_l, 3+atom_nf] draws in the reference's order (z_T, then per iteration
        reverse / known / re-noise, then the final head).  Returns (xh_lig, xh_pocket, lig_mask, pocket_mask)."""
        timesteps = self.T if timesteps is None else timesteps
        dev = self.device
        B = len(ligand['size'])
        nv0, nv1, nb1 = self.norm_values[0], self.norm_values[1], self.norm_biases[1]
        lig_mask = ligand['mask'].to(dev).long()
        pocket_mask = pocket['mask'].to(dev).long()
        lx = ligand['x'].to(dev, torch.float32) / nv0
        lh = (ligand['one_hot'].to(dev).float() - nb1) / nv1
        xh0_pocket = torch.cat([pocket['x'].to(dev, torch.float32) / nv0,
                                (pocket['one_hot'].to(dev).float() - nb1) / nv1], dim=1).contiguous()
        fixed = lig_fixed.to(dev).reshape(-1).float()
        fx = fixed > 0
        n_l = int(lig_mask.numel())
        provider = noise if isinstance(noise, NoiseProvider) else None
        draws = iter(noise) if (noise is not None and provider is None) else None

        def draw_one():
            if provider is not None:
                return self._h2d(torch.as_tensor(provider.step(n_l), dtype=torch.float32))
            return None if draws is None else next(draws)
        nxt = lambda: self._noise(n_l, draw_one())

        def seg_mean(x, idx):
            return segment_mean_sorted(x, idx, B)

        com_pocket_0 = seg_mean(xh0_pocket[:, :3], pocket_mask)
        if center == 'ligand':
            mean_known = seg_mean(lx[fx], lig_mask[fx])
        elif center == 'pocket':
            mean_known = com_pocket_0
        else:
            raise NotImplementedError(f"Centering option {center} not implemented")
        mu = torch.cat([mean_known, torch.zeros((B, self.atom_nf), device=dev)], dim=1)[lig_mask].contiguous()
        ident = torch.tensor([[1.0, 0.0, 1.0]], device=dev).repeat(B, 1)
        z_lig, xh_pocket = self.engine.sampler_step(mu, None, nxt(), xh0_pocket, ident, lig_mask, pocket_mask, B)
        xh_ligand = torch.cat([lx, lh], dim=1).contiguous()
        zero = torch.zeros(B)
        for s in reversed(range(0, timesteps)):
            s_array = torch.full((B, 1), fill_value=s, dtype=torch.float32) / timesteps
            t_array = torch.full((B, 1), fill_value=s + 1, dtype=torch.float32) / timesteps
            g_s, g_t = self.lookup(s_array), self.lookup(t_array)
            alpha_s, sigma_s = torch.sqrt(torch.sigmoid(-g_s)), torch.sqrt(torch.sigmoid(g_s))
            coef_known = torch.stack([alpha_s, zero, sigma_s], dim=1).to(dev)
            sigma2_ts = -torch.expm1(F.softplus(g_s) - F.softplus(g_t))
            alpha_ts = torch.exp(0.5 * (F.logsigmoid(-g_t) - F.logsigmoid(-g_s)))
            coef_renoise = torch.stack([alpha_ts, zero, torch.sqrt(sigma2_ts)], dim=1).to(dev)
            for u in range(resamplings):
                z_unknown, xh_pocket = self.sample_p_zs_given_zt(s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask,
                                                                 noise=draw_one(), n_samples=B)
                if reward_fn is not None and spsa_window[0] <= s <= spsa_window[1] and u < 1:
                    zeta = 1e-3 * (s / 1200)                                              # :1571-1572
                    pn, xn = self._spsa_draws(provider, spsa_k, lig_mask, B)
                    z_upd, xh_pocket = self.my_update_z_lig(z_lig, xh_pocket, lig_mask, pocket_mask, t_array, B, zeta,
                                                            reward_fn, guidance_scale=1e-3, k=spsa_k, perturbation_noise=pn,
                                                            x0_noise=xn)
                    z_unknown, xh_pocket = self._unnormalize_quirk(z_upd, xh_pocket, lig_mask, pocket_mask, B)
                # known atoms follow the pocket's accumulated translation, then q(z_s | x)
                com_pocket = seg_mean(xh_pocket[:, :3], pocket_mask)
                xh_ligand[:, :3] = lx + (com_pocket - com_pocket_0)[lig_mask]
                z_known, xh_pocket = self.engine.sampler_step(xh_ligand, None, nxt(), xh_pocket, coef_known, lig_mask,
                                                              pocket_mask, B)
                dx = seg_mean(z_unknown[fx][:, :3], lig_mask[fx]) - seg_mean(z_known[fx][:, :3], lig_mask[fx])
                z_known[:, :3] += dx[lig_mask]
                xh_pocket[:, :3] += dx[pocket_mask]
                z_lig = (z_known * fixed[:, None] + z_unknown * (1 - fixed[:, None])).contiguous()
                if u < resamplings - 1:
                    z_lig, xh_pocket = self.engine.sampler_step(z_lig, None, nxt(), xh_pocket, coef_renoise, lig_mask,
                                                                pocket_mask, B)
            if svdd == 1 and reward_fn is not None and s <= svdd_schedule[0] and s % svdd_schedule[1] == 0:
                sizes = torch.bincount(lig_mask, minlength=B)
                if int(sizes.min()) != int(sizes.max()):
                    raise NotImplementedError("ATP re-batching inside inpaint keeps ligand['mask'] (conditional_model.py:"
                                              "1626, 1779): only defined for equally sized ligands")
                z_lig, xh_pocket, _ = self._atp_event(s, s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask, B,
                                                      reward_fn, svdd_groups, provider=provider, x0_pocket_group0=xh0_pocket)
                z_lig, xh_pocket = self._unnormalize_quirk(z_lig, xh_pocket, lig_mask, pocket_mask, B)
        x_lig, h_lig, x_pocket, h_pocket = self.sample_p_xh_given_z0(
            z_lig, xh_pocket, lig_mask, pocket_mask, B, noise=draw_one())
        self._raise_on_flags()
        return torch.cat([x_lig, h_lig.float()], dim=1), torch.cat([x_pocket, h_pocket], dim=1), lig_mask, pocket_mask

    # -- ATP ("SVDD") event, conditional_model.py:1085-1241 -------------------------------------------------------------
    def _rebatch(self, top_idx, big_z, big_p, big_lig_mask, n_p):
        """Winners in rank order (:1212-1232) as three gathers (no per-winner Python loop; one sync for the new length)."""
        dev = self.device
        n_cand = int(big_p.shape[0] // n_p)
        sizes = torch.bincount(big_lig_mask, minlength=n_cand)
        starts = torch.cumsum(sizes, 0) - sizes
        sel_sizes = sizes[top_idx]
        total = int(sel_sizes.sum())
        new_m = torch.repeat_interleave(torch.arange(len(top_idx), device=dev), sel_sizes, output_size=total)
        new_starts = torch.cumsum(sel_sizes, 0) - sel_sizes
        src = starts[top_idx][new_m] + (torch.arange(total, device=dev) - new_starts[new_m])
        p_src = (top_idx[:, None] * n_p + torch.arange(n_p, device=dev)[None, :]).reshape(-1)
        return big_z[src].contiguous(), big_p[p_src].contiguous(), new_m

    def _select(self, s, big_z, big_p, x0_l, h0_types, big_lig_mask, B, n_p, reward_fn, pending_r=None):
        """mixed = r0 * (s / 250) + r * (250 - s / 250) [sic, :1203]; global top-B; re-batching."""
        dev = self.device
        reward_fn = _event_fn(reward_fn, 'svdd')
        r0 = torch.as_tensor(reward_fn(x0_l, h0_types, big_lig_mask), dtype=torch.float32, device=dev)
        r = torch.as_tensor(pending_r.result() if pending_r is not None else
                            reward_fn(big_z[:, :3], big_z[:, 3:].argmax(1), big_lig_mask), dtype=torch.float32, device=dev)
        mixed = r0 * (s / 250) + r * (250 - s / 250)
        self.last_atp_rewards = (r0, r)
        _, top_idx = mixed.topk(k=B, largest=True)                                     # :1205
        return self._rebatch(top_idx, big_z, big_p, big_lig_mask, n_p)

    def _atp_event(self, s, s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask, B, reward_fn, n_groups,
                   provider=None, x0_pocket_group0=None):
        """Draw n_groups-1 extra candidate next-states from (z, s, t), score the current and x0-look-ahead molecules of
        all n_groups*B candidates, keep the global top-B (:1203-1232).  The candidate groups are evaluated as ONE batch of
        n_groups*B samples per denoiser call instead of the reference's sequential calls.

        With ``self.atp_group`` set (a torch.distributed group whose ranks carry the SAME trajectory state: same seed of the
        default generator, see ``set_distributed_atp``), the candidate groups are split over the ranks -- group g is drawn
        (from the per-rank candidate generator), denoised and scored on rank g % world -- and the winners are rebuilt
        everywhere from ONE all-gather of scores, latents and absolute pocket positions (parallel.py).

        ``x0_pocket_group0``: pocket used for the x0 look-ahead of the CURRENT state only.  The inpainting loop passes the
        original, untranslated pocket there, as the reference does (``my_to_x0(t_array, z_lig, xh0_pocket, ...)``,
        conditional_model.py:1631, while the extra candidates look ahead from their own translated pockets, :1651)."""
        dev = self.device
        n_l, n_p = z_lig.shape[0], xh_pocket.shape[0]
        G = n_groups
        all_events_fn, reward_fn = reward_fn, _event_fn(reward_fn, 'svdd')
        group = getattr(self, 'atp_group', None)
        world = dist.get_world_size(group) if group is not None else 1
        rank = dist.get_rank(group) if group is not None else 0
        n_extra = len([g for g in range(1, G) if g % world == rank])         # extra groups drawn here
        n_here = n_extra + (1 if rank == 0 else 0)                          # rank 0 also owns the current state (group 0)
        rep = lambda a, n: a.unsqueeze(0).repeat(n, 1, 1).reshape(n * a.shape[0], -1)
        offs = torch.arange(max(n_here, 1), device=dev) * B
        big_lig_mask = (lig_mask.unsqueeze(0) + offs[:, None]).reshape(-1)[:n_here * n_l]
        big_pocket_mask = (pocket_mask.unsqueeze(0) + offs[:, None]).reshape(-1)[:n_here * n_p]
        step_nz = x0_nz = None
        if provider is not None:
            assert world == 1, 'injected draws are defined for the single-rank event'
            step_nz, x0_nz = provider.atp(G, n_l)
            step_nz = self._h2d(torch.as_tensor(step_nz, dtype=torch.float32)).reshape(n_extra * n_l, -1)
            x0_nz = self._h2d(torch.as_tensor(x0_nz, dtype=torch.float32)).reshape(G * n_l, -1)
        gen = self.cand_gen
        if step_nz is None and gen is not None and n_extra > 0:
            step_nz = torch.randn((n_extra * n_l, self.n_dims + self.atom_nf), device=dev, generator=gen)
        if x0_nz is None and gen is not None and n_here > 0:
            x0_nz = torch.randn((n_here * n_l, self.n_dims + self.atom_nf), device=dev, generator=gen)
        parts_z, parts_p = ([z_lig], [xh_pocket]) if rank == 0 else ([], [])
        if n_extra > 0:                # extra candidates: sample_p_zs_given_zt from the same (already denoised) state, :1109-1117
            zs_extra, xp_extra = self.sample_p_zs_given_zt(
                s_array.repeat(n_extra, 1), t_array.repeat(n_extra, 1), rep(z_lig, n_extra), rep(xh_pocket, n_extra),
                big_lig_mask[:n_extra * n_l], big_pocket_mask[:n_extra * n_p], noise=step_nz, n_samples=n_extra * B)
            parts_z.append(zs_extra)
            parts_p.append(xp_extra)
        if n_here > 0:
            big_z = torch.cat(parts_z, dim=0)
            big_p = torch.cat(parts_p, dim=0)
            pending_r = None
            overlap = hasattr(reward_fn, 'submit') and self.overlap_scoring
            if overlap:
                # the current candidates are scored by the worker processes while the GPU runs their x0 look-ahead
                cand_x, cand_t = big_z[:, :3].contiguous(), big_z[:, 3:].argmax(1)
                ready = torch.cuda.Event()
                ready.record()
            look_p = big_p
            if x0_pocket_group0 is not None and rank == 0:
                look_p = torch.cat([x0_pocket_group0.to(big_p.dtype), big_p[n_p:]], dim=0)
            x0_l, h0_l, _, _ = self.my_to_x0(t_array.repeat(n_here, 1), big_z, look_p, big_lig_mask, big_pocket_mask, n_here * B,
                                             noise=x0_nz)
            if overlap:
                pending_r = reward_fn.submit(cand_x, cand_t, big_lig_mask, after=ready)
            if world == 1:
                return self._select(s, big_z, big_p, x0_l, h0_l.argmax(1), big_lig_mask, B, n_p // B, reward_fn, pending_r)
            r0 = torch.as_tensor(reward_fn(x0_l, h0_l.argmax(1), big_lig_mask), dtype=torch.float32, device=dev)
            r = torch.as_tensor(pending_r.result() if pending_r is not None else
                                reward_fn(big_z[:, :3], big_z[:, 3:].argmax(1), big_lig_mask), dtype=torch.float32, device=dev)
            mixed = r0 * (s / 250) + r * (250 - s / 250)                              # [sic] :1203
        else:
            big_z = z_lig[:0]
            big_p = xh_pocket[:0]
            mixed = torch.zeros(0, device=dev)
        from .parallel import atp_select_packed
        return atp_select_packed(mixed, big_z, big_p, lig_mask, pocket_mask, xh_pocket, B, G, group)

    # -- the s == 30 branch of a run with SPSA, conditional_model.py:1261-1418 ---------------------------------------------
    def _mixed_event(self, s, s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask, B, reward_fn, n_groups, zeta,
                     guidance_scale, k, provider=None):
        """Four extra candidates, each one reverse step + SPSA update + feature rescaling; then the ATP selection over the
        entry state and the four.  The reference REBINDS ``z_lig`` / ``xh_pocket`` inside its loop (:1286): candidate i + 1
        is drawn from candidate i's SPSA output (before the rescaling), so the candidates form a chain and cannot be batched
        across groups; ``zeta`` is reset to 1e-3 from the third candidate on (:1284-1285).  The 2k look-aheads inside each
        SPSA update are batched as everywhere else."""
        dev = self.device
        n_l, n_p = z_lig.shape[0], xh_pocket.shape[0]
        G = n_groups
        # With ranks sharing one trajectory (atp_group) every rank computes the whole chain from the shared default
        # generator: the chain cannot be split, and all ranks must leave the event with the same state.
        step_nz = spsa_nz = x0_nz = None
        if provider is not None:
            step_nz, spsa_nz, x0_nz = provider.mixed(G, k, torch.bincount(lig_mask, minlength=B).tolist())
        cands_z, cands_p = [z_lig], [xh_pocket]
        z_cur, xp_cur = z_lig, xh_pocket
        for i in range(G - 1):
            nz = None if step_nz is None else self._h2d(torch.as_tensor(step_nz[i], dtype=torch.float32))
            z_tmp, xp_tmp = self.sample_p_zs_given_zt(s_array, t_array, z_cur, xp_cur, lig_mask, pocket_mask, noise=nz,
                                                      n_samples=B)                               # :1278-1283
            if i >= 2:
                zeta = 1e-3
            pn, xn = spsa_nz[i] if spsa_nz is not None else (None, None)
            z_cur, xp_cur = self.my_update_z_lig(z_tmp, xp_tmp, lig_mask, pocket_mask, t_array, B, zeta, reward_fn,
                                                 guidance_scale=guidance_scale, k=k, perturbation_noise=pn, x0_noise=xn)   # :1286
            z_tmp, xp_tmp = self._unnormalize_quirk(z_cur, xp_cur, lig_mask, pocket_mask, B)  # :1287-1292
            cands_z.append(z_tmp)
            cands_p.append(xp_tmp)
        offs = torch.arange(G, device=dev) * B
        big_lig_mask = (lig_mask.unsqueeze(0) + offs[:, None]).reshape(-1)
        big_pocket_mask = (pocket_mask.unsqueeze(0) + offs[:, None]).reshape(-1)
        big_z = torch.cat(cands_z, dim=0)
        big_p = torch.cat(cands_p, dim=0)
        x0n = None if x0_nz is None else self._h2d(torch.as_tensor(x0_nz, dtype=torch.float32)).reshape(G * n_l, -1)
        x0_l, h0_l, _, _ = self.my_to_x0(t_array.repeat(G, 1), big_z, big_p, big_lig_mask, big_pocket_mask, G * B, noise=x0n)
        return self._select(s, big_z, big_p, x0_l, h0_l.argmax(1), big_lig_mask, B, n_p // B, reward_fn)
