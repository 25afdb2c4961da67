"""CPU restatement (numpy) of the reference's guidance control flow around the denoiser.  TEST INFRASTRUCTURE ONLY: imported
by tests/ (never by the product path, see tests/test_contracts.py).

Follows, call by call and draw by draw, ``ConditionalDDPM`` of /root/reference/equivariant_diffusion/conditional_model.py:

* ``my_to_x0``                         :457-468 (+ sample_p_xh_given_z0 :136-160)
* ``my_perturbation_for_molecule``     :724-736
* ``my_gradient_for_molecule``         :738-759
* ``my_update_z_lig``                  :760-813   (SPSA; k = 10 sequential +/- look-aheads)
* the ATP block of sample_given_pocket :1085-1241 (5 candidate groups, mixed reward, global top-B, re-batching)
* the ``s == 30`` branch               :1261-1418 (4 chained extra candidates, each SPSA-updated; zeta reset to 1e-3 for i >= 2)
* the feature rescaling after an event :1235-1240, 1253-1258 ("unnormalize" applied to the latent, reproduced as is)
* ``handle_to_mol``'s translation      :845-864  (applied to what the reward sees)

Pinned by tests/golden/guidance.npz, which tests/golden/make_golden_guidance.py generates by running the unmodified
reference with stubbed chemistry (tests/test_guidance_oracle.py replays it).

Everything that is external in the reference is injected: ``dyn(z, xp, t[B,1], lig_mask, pocket_mask) -> eps_lig`` (the
denoiser), ``reward_fn(x, types, lig_mask) -> list[float]`` (handle_to_mol + my_reward_for_SPSA/_SVDD) and
``draw(shape) -> float32 array`` (torch.randn), called in the reference's order.
"""
from __future__ import annotations

import numpy as np

from . import egnn_oracle as O

F32 = np.float32


class GuidanceOracle:
    def __init__(self, dyn, reward_fn, draw, pocket_com_before, cfg: O.OracleConfig = O.OracleConfig(), T: int = 500):
        self.dyn, self.reward_fn, self.draw, self.cfg, self.T = dyn, reward_fn, draw, cfg, T
        self.gam = O.gamma_table(T, cfg.noise_precision, 2.0)
        self.com_before = np.asarray(pocket_com_before, F32)          # [B,3]
        self.trace = []                                               # (tag, payload) for tests

    def gamma(self, t):
        return self.gam[np.round(np.asarray(t, F32).reshape(-1) * self.T).astype(np.int64)]

    # -- my_to_x0 :457-468 --------------------------------------------------------------------------------------------
    def my_to_x0(self, t, z, xp, lm, pm):
        nb = len(t)
        eps_t = self.dyn(z, xp, t, lm, pm)
        z0 = O.x0_lookahead_z0(z, eps_t, self.gamma(t), lm)
        t0 = np.zeros((nb, 1), F32)
        eps0 = self.dyn(z0, xp, t0, lm, pm)
        noise = self.draw((len(lm), z.shape[1]))
        x_l, types, x_p, h_p = O.sample_p_xh_given_z0(z0, xp, eps0, noise, self.gamma(t0), lm, pm, self.cfg)
        return x_l, types, x_p, h_p

    # -- handle_to_mol :845-864 + reward ---------------------------------------------------------------------------------
    def score(self, x_lig, types, x_pocket, lm, pm, com_before=None):
        cb = self.com_before if com_before is None else com_before
        nb = len(cb)
        com_after = O.segment_mean(np.asarray(x_pocket, F32)[:, :3], pm, nb).astype(F32)
        x = (np.asarray(x_lig, F32) + (cb - com_after)[lm]).astype(F32)
        self.trace.append(('mol', x.copy(), np.asarray(types).copy()))
        return np.asarray(self.reward_fn(x, types, lm), np.float64)

    # -- SPSA :724-813 -------------------------------------------------------------------------------------------------------
    def my_update_z_lig(self, z, xp, lm, pm, t, zeta, guidance_scale, k=10):
        z = np.asarray(z, F32)
        nb = len(t)
        perts, f_plus, f_minus = [], [], []
        for _ in range(k):
            pert = np.zeros_like(z[:, :3])
            for b in np.unique(lm):                                       # :771-782, one draw per molecule
                idx = np.nonzero(lm == b)[0]
                noise = self.draw((len(idx), 3))
                pert[idx] = F32(zeta) * (noise - noise.mean(axis=0, keepdims=True, dtype=F32))
            zp, zm = z.copy(), z.copy()
            zp[:, :3] = z[:, :3] + pert
            zm[:, :3] = z[:, :3] - pert
            xl, tl, xpk, _ = self.my_to_x0(t, zp, xp, lm, pm)             # :786
            xl_m, tl_m, xpk_m, _ = self.my_to_x0(t, zm, xp, lm, pm)       # :790
            fp = self.score(xl, tl, xpk, lm, pm)                          # :739-743: both molecules first, then both rewards
            fm = self.score(xl_m, tl_m, xpk_m, lm, pm)
            perts.append(pert); f_plus.append(fp); f_minus.append(fm)
        self.trace.append(('spsa_rewards', np.stack(f_plus), np.stack(f_minus)))
        return O.spsa_update(z, xp, np.stack(perts), np.stack(f_plus).astype(F32), np.stack(f_minus).astype(F32), lm, pm,
                             guidance_scale=guidance_scale)

    # -- "unnormalize" of the latent after an event :1235-1240, 1253-1258 -----------------------------------------------------
    def rescale(self, z, xp, lm, pm, nb):
        nv0, nv1, nb1 = F32(self.cfg.norm_values[0]), F32(self.cfg.norm_values[1]), F32(self.cfg.norm_biases[1])
        zx, zh = z[:, :3] * nv0, z[:, 3:] * nv1 + nb1
        px, ph = xp[:, :3] * nv0, xp[:, 3:] * nv1 + nb1
        zx, px = O.remove_mean_batch(zx, px, lm, pm, nb)
        return np.concatenate([zx, zh], 1).astype(F32), np.concatenate([px, ph], 1).astype(F32)

    def reverse_step(self, s, t, z, xp, lm, pm):
        """sample_p_zs_given_zt :483-540 (optimize = 0)."""
        eps = self.dyn(z, xp, t, lm, pm)
        noise = self.draw((len(lm), z.shape[1]))
        return O.sample_p_zs_given_zt(z, xp, eps, noise, self.gamma(s), self.gamma(t), lm, pm)

    # -- candidate selection shared by the ATP block and the s == 30 branch :1129-1240 ----------------------------------------
    def _select(self, s, cands, cands0, lm, pm, nb):
        """cands / cands0: lists of (z, xp) / (x0 with one-hot features, x0 pocket) per group, group 0 = current state."""
        G = len(cands)
        n_l, n_p = len(lm), len(pm)
        big_lm = np.concatenate([lm + g * nb for g in range(G)])          # the reference offsets by i * 20 (= B)
        big_pm = np.concatenate([pm + g * nb for g in range(G)])
        big_cb = np.tile(self.com_before, (G, 1))
        big_z = np.concatenate([c[0] for c in cands]); big_p = np.concatenate([c[1] for c in cands])
        z0 = np.concatenate([c[0] for c in cands0]); p0 = np.concatenate([c[1] for c in cands0])
        r0 = self.score(z0[:, :3], z0[:, 3:].argmax(1), p0, big_lm, big_pm, big_cb)          # :1180-1187
        # handle_to_mol translates big_z / big_p IN PLACE (:855-858); the winners are cut from the translated tensors
        com_after = O.segment_mean(big_p[:, :3], big_pm, G * nb).astype(F32)
        shift = (big_cb - com_after).astype(F32)
        big_p = big_p.copy(); big_z = big_z.copy()
        big_p[:, :3] = big_p[:, :3] + shift[big_pm]
        big_z[:, :3] = big_z[:, :3] + shift[big_lm]
        self.trace.append(('mol', big_z[:, :3].copy(), big_z[:, 3:].argmax(1)))
        r = np.asarray(self.reward_fn(big_z[:, :3], big_z[:, 3:].argmax(1), big_lm), np.float64)   # :1191-1201
        self.trace.append(('atp_rewards', r0, r))
        z_new, p_new, lm_new, order = O.atp_select(r0, r, s, big_z, big_p, big_lm, big_pm, nb)
        self.trace.append(('atp_order', order))
        return self.rescale(z_new, p_new, lm_new, pm, nb) + (lm_new,)

    def _x0_pair(self, t, z, xp, lm, pm):
        xl, tl, xpk, hpk = self.my_to_x0(t, z, xp, lm, pm)
        onehot = np.eye(z.shape[1] - 3, dtype=F32)[tl]
        return np.concatenate([xl, onehot], 1), np.concatenate([xpk, hpk], 1)

    def atp_event(self, s, s_arr, t_arr, z, xp, lm, pm, n_extra=4, x0_pocket_group0=None):
        """:1085-1241; the inpainting loop's copy (:1629-1778) looks ahead from the current state with the ORIGINAL pocket
        ``xh0_pocket`` (:1631) -- pass it as ``x0_pocket_group0``."""
        nb = len(t_arr)
        cands0 = [self._x0_pair(t_arr, z, xp if x0_pocket_group0 is None else x0_pocket_group0, lm, pm)]      # :1095 / :1631
        cands = [(z, xp)]
        for _ in range(n_extra):
            z_tmp, xp_tmp = self.reverse_step(s_arr, t_arr, z, xp, lm, pm)            # :1112-1117
            cands0.append(self._x0_pair(t_arr, z_tmp, xp_tmp, lm, pm))    # :1118
            cands.append((z_tmp, xp_tmp))
            self.trace.append(('cand', z_tmp.copy()))
        return self._select(s, cands, cands0, lm, pm, nb)

    def mixed_event(self, s, s_arr, t_arr, z, xp, lm, pm, zeta, guidance_scale, n_extra=4):
        """:1261-1418.  ``z_lig`` / ``xh_pocket`` are REBOUND inside the loop (:1286): candidate i+1 is drawn from candidate
        i's SPSA output (before its rescaling), while group 0 stays the state the branch was entered with."""
        nb = len(t_arr)
        cands0 = [self._x0_pair(t_arr, z, xp, lm, pm)]                    # :1262
        cands = [(z, xp)]
        z_cur, xp_cur = z, xp
        for i in range(n_extra):
            z_tmp, xp_tmp = self.reverse_step(s_arr, t_arr, z_cur, xp_cur, lm, pm)    # :1278-1283
            if i >= 2:
                zeta = 1e-3                                               # :1284-1285
            z_cur, xp_cur = self.my_update_z_lig(z_tmp, xp_tmp, lm, pm, t_arr, zeta, guidance_scale)     # :1286
            z_tmp, xp_tmp = self.rescale(z_cur, xp_cur, lm, pm, nb)       # :1287-1292
            cands0.append(self._x0_pair(t_arr, z_tmp, xp_tmp, lm, pm))    # :1294
            cands.append((z_tmp, xp_tmp))
            self.trace.append(('cand', z_tmp.copy()))
        return self._select(s, cands, cands0, lm, pm, nb)

    # -- the sampling loop :944-1420 from a given state ----------------------------------------------------------------------------
    def run(self, z, xp, lm, pm, s_from, s_to, timesteps, svdd, spsa, on_state=None):
        nb = int(pm.max()) + 1
        for s in range(s_from, s_to - 1, -1):
            s_arr = np.full((nb, 1), s, F32) / F32(timesteps)
            t_arr = (np.full((nb, 1), s, F32) + F32(1)) / F32(timesteps)
            z, xp = self.reverse_step(s_arr, t_arr, z, xp, lm, pm)
            if svdd == 1 and s <= 50 and s % 10 == 0:
                z, xp, lm = self.atp_event(s, s_arr, t_arr, z, xp, lm, pm)
            if spsa == 1 and s <= 30 and s % 2 == 0:
                zeta = 1e-3 * (s / 500)
                z, xp = self.my_update_z_lig(z, xp, lm, pm, t_arr, zeta, 1e-3)
                z, xp = self.rescale(z, xp, lm, pm, nb)
                if s == 30:
                    z, xp, lm = self.mixed_event(s, s_arr, t_arr, z, xp, lm, pm, zeta, 1e-3)
            if on_state is not None:
                on_state(s, z, xp, lm)
        return z, xp, lm
