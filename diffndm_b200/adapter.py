"""Call-compatible stand-in for the reference's ``ConditionalDDPM`` on the sampling path.

``LigandPocketDDPM.generate_ligands`` (lightning_modules.py:899-901) calls ``self.ddpm.sample_given_pocket`` with the
reference's 14 positional arguments and ``self.ddpm.inpaint`` with its own list (conditional_model.py:886-887, 1492-1493);
swapping only ``model.ddpm.dynamics`` keeps the reference's Python loop around the engine -- one ``.item()`` and
``empty_cache()`` per step, 40 sequential look-aheads per SPSA update.  ``B200ConditionalDDPM`` takes the same arguments,
runs the whole loop on the engine (CUDA-graph replay of the unguided stretch, guidance copies batched) and returns what the
reference returns, so that

    model.ddpm = B200ConditionalDDPM.from_reference(model.ddpm)

is the drop-in.  The host-chemistry arguments (``dataset_info, sanitize, relax_iter, largest_frag``) configure the reward
of the guided modes (``rewards_rdkit``, needs RDKit / OpenBabel on the host) unless a ``reward_fn`` was given; the RL noise
adjustment (``optimize = 1``, AdjustNet, conditional_model.py:509-518, 1440-1484) is training-side and out of scope.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .engine import B200EGNNDynamics
from .sampler import ConditionalSampler


class B200ConditionalDDPM(torch.nn.Module):
    def __init__(self, dynamics: B200EGNNDynamics, timesteps: int = 500, noise_schedule: str = 'polynomial_2',
                 noise_precision: float = 5.0e-4, norm_values=(1.0, 4.0), norm_biases=(None, 0.0),
                 reward_fn: Optional[Callable] = None, reward_workers: int = 8, gamma_table=None):
        super().__init__()
        self.dynamics = dynamics
        self.sampler = ConditionalSampler(dynamics, timesteps=timesteps, noise_schedule=noise_schedule,
                                          noise_precision=noise_precision, norm_values=tuple(norm_values),
                                          norm_biases=tuple(norm_biases), gamma_table=gamma_table)
        self.T = timesteps
        self.n_dims = 3
        self.atom_nf = dynamics.cfg.atom_nf
        self.residue_nf = dynamics.cfg.residue_nf
        self.norm_values, self.norm_biases = tuple(norm_values), tuple(norm_biases)
        self.reward_fn = reward_fn
        self.reward_workers = reward_workers

    @classmethod
    def from_reference(cls, ddpm, reward_fn: Optional[Callable] = None, **engine_kw) -> 'B200ConditionalDDPM':
        """Build from a live reference ``ConditionalDDPM`` (weights of ``ddpm.dynamics`` are packed once)."""
        dyn = ddpm.dynamics if isinstance(ddpm.dynamics, B200EGNNDynamics) else B200EGNNDynamics.from_reference(ddpm.dynamics, **engine_kw)
        table = ddpm.gamma.gamma if hasattr(getattr(ddpm, 'gamma', None), 'gamma') else None   # the schedule's lookup table itself
        return cls(dyn.eval(), timesteps=int(ddpm.T), norm_values=tuple(ddpm.norm_values), norm_biases=tuple(ddpm.norm_biases),
                   reward_fn=reward_fn, gamma_table=table)

    # -- pieces callers use directly --------------------------------------------------------------------------------------
    def sample_p_zs_given_zt(self, s, t, zt_lig, xh0_pocket, ligand_mask, pocket_mask, optimize=0, fix_noise=False):
        """conditional_model.py:483-540: returns (zs_lig, xh_pocket, log_prob_adjust) like the reference.  The third value
        belongs to the RL noise adjustment, which does not run at ``optimize = 0``: a zero scalar."""
        if optimize == 1:
            raise NotImplementedError('optimize=1 (AdjustNet noise adjustment, RL fine-tuning) is outside the sampling path')
        if fix_noise:
            raise NotImplementedError("fix_noise option isn't implemented yet")          # as the reference, :170-172
        zs, xp = self.sampler.sample_p_zs_given_zt(s, t, zt_lig, xh0_pocket, ligand_mask, pocket_mask)
        return zs, xp, torch.zeros((), device=zs.device)

    def sample_p_xh_given_z0(self, z0_lig, xh0_pocket, lig_mask, pocket_mask, batch_size, fix_noise=False):
        return self.sampler.sample_p_xh_given_z0(z0_lig, xh0_pocket, lig_mask, pocket_mask, batch_size)

    def my_to_x0(self, t, zt_lig, xh0_pocket, ligand_mask, pocket_mask, n_samples):
        return self.sampler.my_to_x0(t, zt_lig, xh0_pocket, ligand_mask, pocket_mask, n_samples)

    def _rewards(self, svdd, spsa, sanitize, relax_iter, largest_frag):
        if not (svdd == 1 or spsa == 1):
            return None
        if self.reward_fn is not None:
            return self.reward_fn
        from .rewards_rdkit import GuidanceRewards, _require
        _require()
        return GuidanceRewards(sanitize=sanitize, relax_iter=relax_iter, largest_frag=largest_frag, workers=self.reward_workers)

    # -- the two entry points of generate_ligands ---------------------------------------------------------------------------
    @torch.no_grad()
    def sample_given_pocket(self, pocket, num_nodes_lig, pocket_com_before=None, dataset_info=None, sanitize=False,
                            relax_iter=0, largest_frag=False, pdb_id=None, device=None, optimize=0, path=None,
                            path_save=None, svdd=0, spsa=0, return_frames=1, timesteps=None):
        """conditional_model.py:886-887 -> (xh_lig, xh_pocket, lig_mask, pocket_mask), :1488-1489."""
        if optimize == 1:
            raise NotImplementedError('optimize=1 (AdjustNet noise adjustment, RL fine-tuning) is outside the sampling path')
        timesteps = self.T if timesteps is None else timesteps
        assert 0 < return_frames <= timesteps and timesteps % return_frames == 0            # :906-907
        if return_frames != 1:
            raise NotImplementedError('intermediate frames are not kept (the reference writes only frame 0 as well, :1440-1443)')
        rewards = self._rewards(svdd, spsa, sanitize, relax_iter, largest_frag)
        own = rewards is not None and rewards is not self.reward_fn
        try:
            return self.sampler.sample_given_pocket(pocket, num_nodes_lig, timesteps=timesteps, svdd=int(svdd), spsa=int(spsa),
                                                    reward_fn=rewards)
        finally:
            if own:
                rewards.close()

    @torch.no_grad()
    def inpaint(self, ligand, pocket, lig_fixed, svdd=0, pocket_com_before=None, dataset_info=None, sanitize=False,
                relax_iter=0, largest_frag=False, resamplings=1, return_frames=1, timesteps=None, center='ligand'):
        """conditional_model.py:1492-1493 -> (xh_lig, xh_pocket, lig_mask, pocket_mask).  The reference enters its hard-wired
        SPSA window (12 <= s <= 16) on every inpainting run, which needs the host chemistry; without RDKit and without a
        ``reward_fn`` the window is skipped (identical for runs of fewer than 13 steps)."""
        if return_frames != 1:
            raise NotImplementedError('intermediate frames are not kept')
        rewards = self.reward_fn
        own = False
        if rewards is None:
            from .rewards_rdkit import rdkit_available, GuidanceRewards
            if rdkit_available():
                rewards, own = GuidanceRewards(sanitize=sanitize, relax_iter=relax_iter, largest_frag=largest_frag,
                                               workers=self.reward_workers), True
            elif svdd == 1:
                from .rewards_rdkit import _require
                _require()
        try:
            return self.sampler.inpaint(ligand, pocket, lig_fixed, svdd=int(svdd), resamplings=resamplings, timesteps=timesteps,
                                        center=center, reward_fn=rewards)
        finally:
            if own:
                rewards.close()
