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


class _GraphedReverseStep:
    """One unguided reverse step (denoiser forward + noise draw + fused p(z_s|z_t) update, in place on the state buffers)
    captured once in a CUDA graph and replayed with new (t, coefficient) values -- removes the ~60 kernel launches and the
    per-call host work from every step of a trajectory.  Valid while the mask tensors and the batch layout stay the same
    (between ATP events); flags are sticky on the device and read by the caller."""

    def __init__(self, sampler: 'ConditionalSampler', z_lig, xh_pocket, lig_mask, pocket_mask, B):
        eng = sampler.engine
        dev = sampler.device
        self.z, self.xp = z_lig, xh_pocket                     # updated in place
        self.lig_mask, self.pocket_mask = lig_mask, pocket_mask   # the graph reads these buffers on every replay
        self.t_buf = torch.zeros((B, 1), device=dev)
        self.coef_buf = torch.zeros((B, 3), device=dev)

        def body():
            eps, _ = eng.forward(self.z, self.xp, self.t_buf, lig_mask, pocket_mask, B, want_pocket=False)
            nz = torch.randn_like(self.z)
            eng.sampler_step(self.z, eps, nz, self.xp, self.coef_buf, lig_mask, pocket_mask, B, z_out=self.z,
                             pocket_out=self.xp, check_com=True)

        keep_z, keep_p = self.z.clone(), self.xp.clone()
        rng = torch.cuda.get_rng_state(dev)                    # the dry runs below must not consume the caller's noise stream
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()                                             # warm-up outside capture (allocator, layout cache)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()
        self.z.copy_(keep_z)
        self.xp.copy_(keep_p)
        torch.cuda.synchronize(dev)
        torch.cuda.set_rng_state(rng, dev)
        eng.read_flags()                                       # the two dry runs may have tripped the COM-drift flag

    def __call__(self, t_dev: torch.Tensor, coef_dev: torch.Tensor):
        """t_dev: 0-d device tensor, coef_dev: [3] device tensor (same for every sample of an unguided step)."""
        self.t_buf.copy_(t_dev.expand_as(self.t_buf))
        self.coef_buf.copy_(coef_dev.expand_as(self.coef_buf))
        self.graph.replay()


class ConditionalSampler:
    """Sampler over a ``B200EGNNDynamics``; all state lives on the GPU."""

    def __init__(self, dynamics: B200EGNNDynamics, timesteps: int = 500, noise_schedule: str = 'polynomial_2',
                 noise_precision: float = 5.0e-4, norm_values=(1.0, 4.0), norm_biases=(None, 0.0),
                 check_every_step: bool = False):
        assert not dynamics.update_pocket_coords           # conditional_model.py:24
        self.dynamics = dynamics
        self.engine = dynamics.engine
        self.T = timesteps
        self.n_dims = 3
        self.atom_nf = dynamics.cfg.atom_nf
        self.norm_values = norm_values
        self.norm_biases = norm_biases
        self.check_every_step = check_every_step
        self.overlap_scoring = True                # SPSA: score one half of a round on the host while the GPU denoises the other
        self._graph_cache = {}                     # (weights version, B, N_l, N_p) -> _GraphedReverseStep
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
            return self.dynamics(z_lig, xh_pocket, t, lig_mask, pocket_mask, n_samples=B)[0]
        finally:
            self.dynamics.compute_pocket_output = keep

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
        p = torch.zeros((x_pocket.shape[0], 3 + self.atom_nf), device=self.device)
        p[:, :3] = x_pocket
        coef = self._const_rows((1.0, 0.0, 0.0), B)
        zo, po = self.engine.sampler_step(z, None, z, p, coef, lig_mask, pocket_mask, B)
        return zo[:, :3], po[:, :3]

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
                        guidance_scale=1e-3, k=10, perturbations=None, x0_noise=None):
        """Symmetric finite-difference guidance.  The reference runs 2k x my_to_x0 sequentially (:764-800); here the
        2k copies are concatenated along the batch axis: ONE denoiser call at t and ONE at t=0 on 2k*B samples.
        ``perturbations`` [k, N_l, 3] / ``x0_noise`` [2k, N_l, 13] may be injected for parity tests."""
        B, n_l, n_p = int(n_samples), z_lig.shape[0], xh_pocket.shape[0]
        sizes = torch.bincount(lig_mask, minlength=B)
        if perturbations is None:                                                   # my_perturbation_for_molecule :724-736
            noise = torch.randn((k, n_l, 3), device=self.device)
            mean = torch.zeros((k, B, 3), device=self.device).index_add_(1, lig_mask, noise) / sizes[None, :, None]
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
        nz = None if x0_noise is None else x0_noise.reshape(reps * n_l, -1)
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
        dd = (f_plus - f_minus) / (2 * 1e-4)                                         # hard-coded divisor, :799
        grad = (dd[:, lig_mask, None] * U).mean(0)                                   # :749-758, :801
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
                            reward_fn: Optional[Callable] = None, noise: Optional[torch.Tensor] = None,
                            spsa_schedule=(30, 2), svdd_schedule=(50, 10), svdd_groups: int = 5, spsa_k: int = 10,
                            use_cuda_graph: bool = True):
        """ConditionalDDPM.sample_given_pocket (conditional_model.py:886-1489) without the host chemistry arguments.

        ``pocket``: dict with 'x' [N_p,3], 'one_hot' [N_p,residue_nf], 'size' [B], 'mask' [N_p] (prepare_pocket layout,
        lightning_modules.py:763-801).  ``noise`` [timesteps+2, N_l, 13] injects the Gaussian draws in the reference's
        order (z_T, one per step, final head) -- only valid for the unguided path.
        ``use_cuda_graph``: replay the unguided reverse step from a CUDA graph (re-captured after every ATP event, whose
        re-batching changes the masks); the NaN / COM-drift flags are then checked at guidance events and at the end
        instead of at the failing step (``check_every_step=True`` keeps the reference's per-step behaviour, eagerly).
        Returns (xh_lig [N_l, 3+atom_nf] with one-hot features, xh_pocket, lig_mask, pocket_mask) like the reference.
        """
        timesteps = self.T if timesteps is None else timesteps
        dev = self.device
        B = len(pocket['size'])
        x_p = pocket['x'].to(dev, torch.float32) / self.norm_values[0]
        h_p = (pocket['one_hot'].to(dev).float() - self.norm_biases[1]) / self.norm_values[1]
        pocket_mask = pocket['mask'].to(dev).long()
        xh0_pocket = torch.cat([x_p, h_p], dim=1).contiguous()
        sizes = torch.as_tensor(num_nodes_lig, device=dev).long()
        lig_mask = torch.repeat_interleave(torch.arange(B, device=dev), sizes)          # utils.py:145-153
        n_l = int(lig_mask.numel())
        # z_T ~ N(pocket COM, I), projected to the ligand-COM-free subspace (:923-930)
        cnt = torch.bincount(pocket_mask, minlength=B).clamp(min=1).float()
        mu_x = torch.zeros((B, 3), device=dev).index_add_(0, pocket_mask, x_p) / cnt[:, None]
        mu = torch.cat([mu_x, torch.zeros((B, self.atom_nf), device=dev)], dim=1)[lig_mask].contiguous()
        ident = torch.tensor([[1.0, 0.0, 1.0]], device=dev).repeat(B, 1)
        z_lig, xh_pocket = self.engine.sampler_step(mu, None, self._noise(n_l, None if noise is None else noise[0]),
                                                    xh0_pocket, ident, lig_mask, pocket_mask, B)
        step = 0
        self.engine.set_static_masks(True)          # lig_mask / pocket_mask are fixed tensors between ATP events
        graphed = use_cuda_graph and noise is None and not self.check_every_step
        gstep = None
        if graphed:                                 # per-step scalars of the whole trajectory, on the device, once
            s_all = torch.arange(timesteps, dtype=torch.float32)
            coef_all = self.step_coefficients(self.lookup(s_all / timesteps), self.lookup((s_all + 1) / timesteps)).to(dev)
            t_all = ((s_all + 1) / timesteps).to(dev)
        nan_check, self.dynamics.check_nan = self.dynamics.check_nan, (self.dynamics.check_nan and not graphed)
        try:
            for s in reversed(range(0, timesteps)):
                s_array = torch.full((B, 1), fill_value=s, dtype=torch.float32) / timesteps
                t_array = torch.full((B, 1), fill_value=s + 1, dtype=torch.float32) / timesteps
                step += 1
                # inside a guidance window the state is replaced every few steps (and ATP re-batches it): capturing a graph
                # there costs more than the launches it saves, so those steps run eagerly (without per-step host syncs)
                in_window = (svdd == 1 and s <= svdd_schedule[0]) or (spsa == 1 and s <= spsa_schedule[0])
                if graphed and not in_window:
                    if gstep is None:
                        # graphs are kept across trajectories: the same pocket with the same ligand sizes (the usual
                        # "n_samples per pocket" loop) replays the graph captured for the first batch
                        key = (self.engine.weights_version, B, int(z_lig.shape[0]), int(xh_pocket.shape[0]))
                        gstep = self._graph_cache.get(key)
                        if gstep is not None and torch.equal(gstep.lig_mask, lig_mask) and \
                                torch.equal(gstep.pocket_mask, pocket_mask):
                            gstep.z.copy_(z_lig)
                            gstep.xp.copy_(xh_pocket)
                        else:
                            gstep = _GraphedReverseStep(self, z_lig.contiguous().clone(), xh_pocket.contiguous().clone(),
                                                        lig_mask, pocket_mask, B)
                            if len(self._graph_cache) >= 4:
                                self._graph_cache.pop(next(iter(self._graph_cache)))
                            self._graph_cache[key] = gstep
                        z_lig, xh_pocket = gstep.z, gstep.xp
                    gstep(t_all[s], coef_all[s])
                else:
                    z_lig, xh_pocket = self.sample_p_zs_given_zt(s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask,
                                                                 noise=None if noise is None else noise[step], n_samples=B)
                if svdd == 1 and s <= svdd_schedule[0] and s % svdd_schedule[1] == 0:
                    if graphed:
                        self._raise_on_flags()
                    self.engine.set_static_masks(False)     # candidate batches use other masks; the winners get a new one
                    z_lig, xh_pocket, lig_mask = self._atp_event(s, s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask,
                                                                 B, reward_fn, svdd_groups)
                    z_lig, xh_pocket = self._unnormalize_quirk(z_lig, xh_pocket, lig_mask, pocket_mask, B)
                    self.engine.set_static_masks(True)
                    gstep = None                            # new state tensors and ligand mask: capture again
                if spsa == 1 and s <= spsa_schedule[0] and s % spsa_schedule[1] == 0:
                    if graphed:
                        self._raise_on_flags()
                    zeta = 1e-3 * (s / 500)                                                   # :1244-1245
                    z_lig, xh_pocket = self.my_update_z_lig(z_lig, xh_pocket, lig_mask, pocket_mask, t_array, B, zeta,
                                                            reward_fn, guidance_scale=1e-3, k=spsa_k)
                    z_lig, xh_pocket = self._unnormalize_quirk(z_lig, xh_pocket, lig_mask, pocket_mask, B)
                    gstep = None
        finally:
            self.dynamics.check_nan = nan_check
        x_lig, h_lig, x_pocket, h_pocket = self.sample_p_xh_given_z0(
            z_lig, xh_pocket, lig_mask, pocket_mask, B, noise=None if noise is None else noise[timesteps + 1])
        self.engine.set_static_masks(False)
        self._raise_on_flags()
        # CoG drift correction (:1431-1438)
        cog = torch.zeros((B, 3), device=dev).index_add_(0, lig_mask, x_lig).abs().max().item()
        if cog > 5e-2:
            x_lig, x_pocket = self.remove_mean_batch(x_lig, x_pocket, lig_mask, pocket_mask, B)
        return torch.cat([x_lig, h_lig.float()], dim=1), torch.cat([x_pocket, h_pocket], dim=1), lig_mask, pocket_mask

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
        draws = iter(noise) if noise is not None else None
        nxt = lambda: self._noise(n_l, None if draws is None else next(draws))

        def seg_mean(x, idx):
            cnt = torch.bincount(idx, minlength=B).clamp(min=1).float()
            return torch.zeros((B, x.shape[1]), device=dev).index_add_(0, idx, x) / cnt[:, None]

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
                                                                 noise=None if draws is None else next(draws), n_samples=B)
                if reward_fn is not None and spsa_window[0] <= s <= spsa_window[1] and u < 1:
                    zeta = 1e-3 * (s / 1200)                                              # :1571-1572
                    z_upd, xh_pocket = self.my_update_z_lig(z_lig, xh_pocket, lig_mask, pocket_mask, t_array, B, zeta,
                                                            reward_fn, guidance_scale=1e-3, k=spsa_k)
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
                                                      reward_fn, svdd_groups)
                z_lig, xh_pocket = self._unnormalize_quirk(z_lig, xh_pocket, lig_mask, pocket_mask, B)
        x_lig, h_lig, x_pocket, h_pocket = self.sample_p_xh_given_z0(
            z_lig, xh_pocket, lig_mask, pocket_mask, B, noise=None if draws is None else next(draws))
        self._raise_on_flags()
        return torch.cat([x_lig, h_lig.float()], dim=1), torch.cat([x_pocket, h_pocket], dim=1), lig_mask, pocket_mask

    # -- ATP ("SVDD") event, conditional_model.py:1085-1241 -------------------------------------------------------------
    def _atp_event(self, s, s_array, t_array, z_lig, xh_pocket, lig_mask, pocket_mask, B, reward_fn, n_groups):
        """Draw n_groups-1 extra candidate next-states from (z, s, t), score the current and x0-look-ahead molecules of
        all n_groups*B candidates, keep the global top-B (:1203-1232).  The candidate groups are evaluated as ONE batch of
        n_groups*B samples per denoiser call instead of the reference's sequential calls.

        With ``self.atp_group`` set (a torch.distributed group whose ranks carry the SAME trajectory state), the candidate
        groups are split over the ranks -- group g is drawn, denoised and scored on rank g % world -- and the winners are
        rebuilt everywhere from one all-gather of scores, latents and pocket translations (parallel.py)."""
        dev = self.device
        n_l, n_p = z_lig.shape[0], xh_pocket.shape[0]
        G = n_groups
        group = getattr(self, 'atp_group', None)
        world = dist.get_world_size(group) if group is not None else 1
        rank = dist.get_rank(group) if group is not None else 0
        n_extra = len([g for g in range(1, G) if g % world == rank])         # extra groups drawn here
        n_here = n_extra + (1 if rank == 0 else 0)                          # rank 0 also owns the current state (group 0)
        rep = lambda a, n: a.unsqueeze(0).repeat(n, 1, 1).reshape(n * a.shape[0], -1)
        offs = torch.arange(max(n_here, 1), device=dev) * B
        big_lig_mask = (lig_mask.unsqueeze(0) + offs[:, None]).reshape(-1)[:n_here * n_l]
        big_pocket_mask = (pocket_mask.unsqueeze(0) + offs[:, None]).reshape(-1)[:n_here * n_p]
        parts_z, parts_p = ([z_lig], [xh_pocket]) if rank == 0 else ([], [])
        if n_extra > 0:                # extra candidates: sample_p_zs_given_zt from the same (already denoised) state, :1109-1117
            zs_extra, xp_extra = self.sample_p_zs_given_zt(
                s_array.repeat(n_extra, 1), t_array.repeat(n_extra, 1), rep(z_lig, n_extra), rep(xh_pocket, n_extra),
                big_lig_mask[:n_extra * n_l], big_pocket_mask[:n_extra * n_p], n_samples=n_extra * B)
            parts_z.append(zs_extra)
            parts_p.append(xp_extra)
        if n_here > 0:
            big_z = torch.cat(parts_z, dim=0)
            big_p = torch.cat(parts_p, dim=0)
            pending_r = None
            if hasattr(reward_fn, 'submit') and self.overlap_scoring:
                # the current candidates are scored by the worker processes while the GPU runs their x0 look-ahead
                cand_x, cand_t = big_z[:, :3].contiguous(), big_z[:, 3:].argmax(1)
                ready = torch.cuda.Event()
                ready.record()
            x0_l, h0_l, _, _ = self.my_to_x0(t_array.repeat(n_here, 1), big_z, big_p, big_lig_mask, big_pocket_mask, n_here * B)
            if hasattr(reward_fn, 'submit') and self.overlap_scoring:
                pending_r = reward_fn.submit(cand_x, cand_t, big_lig_mask, after=ready)
            r0 = torch.as_tensor(reward_fn(x0_l, h0_l.argmax(1), big_lig_mask), dtype=torch.float32, device=dev)
            r = torch.as_tensor(pending_r.result() if pending_r is not None else
                                reward_fn(big_z[:, :3], big_z[:, 3:].argmax(1), big_lig_mask), dtype=torch.float32, device=dev)
            mixed = r0 * (s / 250) + r * (250 - s / 250)                              # [sic] :1203
        else:
            big_z = z_lig[:0]
            big_p = xh_pocket[:0]
            mixed = torch.zeros(0, device=dev)
        if world > 1:
            from .parallel import atp_select_distributed
            sizes = torch.bincount(lig_mask, minlength=B)
            pstart = torch.searchsorted(pocket_mask, torch.arange(B, device=dev))      # first pocket atom of every sample
            src = torch.arange(B, device=dev).repeat(n_here)                           # source sample of every candidate
            first = (torch.arange(n_here, device=dev) * n_p).repeat_interleave(B) + pstart.repeat(n_here)
            shift = big_p[first, :3] - xh_pocket[pstart.repeat(n_here), :3] if n_here > 0 else big_p[:0, :3]
            payload = torch.cat([shift, src[:, None].float()], dim=1)
            z_sel, m_sel, _, pay = atp_select_distributed(mixed, big_z, sizes.repeat(n_here), B, group=group,
                                                          per_candidate=payload)
            new_p = []
            for k in range(B):                                                        # winners' pockets, rank order
                rows = xh_pocket[pocket_mask == int(pay[k, 3])].clone()
                rows[:, :3] += pay[k, :3]
                new_p.append(rows)
            return z_sel.contiguous(), torch.cat(new_p, 0).contiguous(), m_sel
        _, top_idx = mixed.topk(k=B, largest=True)                                     # :1205
        new_z, new_p, new_m = [], [], []
        for rank_pos, idx in enumerate(top_idx.tolist()):                              # :1212-1227
            nm = big_lig_mask == idx
            new_z.append(big_z[nm])
            new_p.append(big_p[big_pocket_mask == idx])
            new_m.append(torch.full((int(nm.sum()),), rank_pos, dtype=torch.long, device=dev))
        return torch.cat(new_z, 0).contiguous(), torch.cat(new_p, 0).contiguous(), torch.cat(new_m, 0)
