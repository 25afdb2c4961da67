"""CPU: the numpy guidance oracle (oracle/guidance_oracle.py) replays the SPSA / ATP / s == 30 events that
tests/golden/make_golden_guidance.py recorded from the UNMODIFIED reference (teacher-forced: every event starts from the
reference's recorded state, consumes the same Gaussian draws and the same stand-in reward function)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_npz_groups

sys.path.insert(0, GOLDEN)
from guidance_common import NoiseStream, geometric_reward, polynomial2_gamma, score_numpy  # noqa: E402
from oracle import egnn_oracle as O  # noqa: E402
from oracle.guidance_oracle import GuidanceOracle  # noqa: E402

CFG = O.OracleConfig()
CASE = 'mixed_b20'


class Fixture:
    """guidance.npz of one case: inputs, per-event records and the regenerated draw list."""

    def __init__(self, case=CASE):
        z = np.load(os.path.join(GOLDEN, 'guidance.npz'))
        pre = case + '/'
        self.top = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre) and k.count('/') == 1}
        self.ev = {}
        for k in z.files:
            if k.startswith(pre) and k.count('/') == 2:
                _, e, kk = k.split('/')
                self.ev.setdefault(e, {})[kk] = z[k]
        self.meta = {k: z[k] for k in z.files if '/' not in k}
        stream = NoiseStream(int(self.top['noise_seed']))
        self.draws = [stream.draw(*[int(v) for v in s if v > 0]) for s in self.top['draw_shapes']]
        self.index = self.top['event_index']                  # (kind id, s, first draw)
        self.B = len(self.top['sizes'])
        self.n_p = len(self.top['pocket_x'])
        self.pm = np.repeat(np.arange(self.B, dtype=np.int64), self.n_p)

    def d0(self, kind, s):
        kid = ['step', 'atp', 'spsa', 'mixed', 'final'].index(kind)
        rows = self.index[(self.index[:, 0] == kid) & (self.index[:, 1] == s)]
        assert len(rows) == 1
        return int(rows[0, 2])

    def draw_from(self, d0):
        it = iter(range(d0, len(self.draws)))

        def draw(shape):
            a = self.draws[next(it)]
            assert tuple(a.shape) == tuple(shape), (a.shape, shape)
            return a
        return draw


@pytest.fixture(scope='module')
def fx():
    return Fixture()


def make_oracle(fx, W, d0, dtype=np.float32):
    gam = polynomial2_gamma().numpy()

    def dyn(z, xp, t, lm, pm):
        eps, _ = O.dynamics_forward(W, z, xp, np.asarray(t, np.float32), lm, pm, CFG, dtype=dtype)
        return score_numpy(eps.astype(np.float32), z, xp, t, lm, pm, fx.top['x0_rel'], gam)

    return GuidanceOracle(dyn, geometric_reward, fx.draw_from(d0), fx.top['com_before'], CFG)


def _arrs(fx, s):
    B = fx.B
    s_arr = np.full((B, 1), s, np.float32) / np.float32(500)
    t_arr = (np.full((B, 1), s, np.float32) + np.float32(1)) / np.float32(500)
    return s_arr, t_arr


def test_fixture_matches_weights(fx, golden_weights):
    from diffndm_b200.weights import weights_checksum
    assert abs(weights_checksum(golden_weights) - float(fx.meta['weights_checksum'])) < 1e-6 * abs(float(fx.meta['weights_checksum']))
    kinds = fx.index[:, 0]
    assert (kinds == 1).sum() == 6 and (kinds == 2).sum() == 16 and (kinds == 3).sum() == 1      # ATP, SPSA, s == 30


@pytest.mark.parametrize('s', [30, 0])
def test_spsa_event_matches_reference(fx, golden_weights, s):
    e = fx.ev[f'spsa{s}']
    orc = make_oracle(fx, golden_weights, fx.d0('spsa', s))
    _, t_arr = _arrs(fx, s)
    lm = e['lm'].astype(np.int64)
    z, xp = orc.my_update_z_lig(e['z_in'], e['xp_in'], lm, fx.pm, t_arr, float(e['zeta']), float(e['guidance_scale']))
    fp, fm = [t for t in orc.trace if t[0] == 'spsa_rewards'][0][1:]
    mols = np.stack([t[1] for t in orc.trace if t[0] == 'mol'])              # plus_0, minus_0, plus_1, ...
    ref_mols = e['mol_x']                                                   # plus_0..plus_9, minus_0..minus_9
    k = len(fp)
    mine = np.concatenate([mols[0::2], mols[1::2]])
    scale = max(1.0, float(np.abs(e['z_in'][:, 3:]).max()))                 # the reference's own feature blow-up (x4 per event)
    if scale <= 64:
        assert np.abs(mine - ref_mols).max() < 2e-4                         # x0 look-ahead molecules, A
        assert np.abs(fp - e['f_plus']).max() < 2e-4 and np.abs(fm - e['f_minus']).max() < 2e-4
    # the update itself with the reference's rewards: pins :738-759, 799-812 to fp32 rounding
    zr, xpr = O.spsa_update(e['z_in'], e['xp_in'], _perturbations(fx, s, lm, float(e['zeta']), k),
                            e['f_plus'].astype(np.float32), e['f_minus'].astype(np.float32), lm, fx.pm,
                            guidance_scale=float(e['guidance_scale']))
    assert np.abs(zr - e['z_out']).max() < 2e-6 * scale and np.abs(xpr - e['xp_out'])[:, :3].max() < 2e-6
    if scale <= 64:
        assert np.abs(z - e['z_out'])[:, :3].max() < 2e-6
    # the feature rescaling that follows the event (:1253-1258)
    za, xpa = orc.rescale(e['z_out'], e['xp_out'], lm, fx.pm, fx.B)
    assert np.allclose(za, e['z_after'], rtol=1e-6, atol=1e-6) and np.allclose(xpa, e['xp_after'], rtol=1e-6, atol=1e-6)


def _perturbations(fx, s, lm, zeta, k):
    """zeta * (noise - mean) per molecule from the recorded draws of one my_update_z_lig call (:724-736, 771-782)."""
    d = fx.d0('spsa', s)
    nb = int(lm.max()) + 1
    out = []
    for _ in range(k):
        pert = np.zeros((len(lm), 3), np.float32)
        for b in range(nb):
            n = fx.draws[d]; d += 1
            pert[lm == b] = np.float32(zeta) * (n - n.mean(axis=0, keepdims=True, dtype=np.float32))
        d += 2                                                               # the two x0 draws
        out.append(pert)
    return np.stack(out)


@pytest.mark.parametrize('s', [50, 30, 20])
def test_atp_event_matches_reference(fx, golden_weights, s):
    e = fx.ev[f'atp{s}']
    orc = make_oracle(fx, golden_weights, fx.d0('atp', s))
    s_arr, t_arr = _arrs(fx, s)
    lm = e['lm'].astype(np.int64)
    z, xp, lm_new = orc.atp_event(s, s_arr, t_arr, e['z_in'], e['xp_in'], lm, fx.pm)
    scale = max(1.0, float(np.abs(e['z_in'][:, 3:]).max()))
    cands = np.stack([t[1] for t in orc.trace if t[0] == 'cand'])
    r0, r = [t for t in orc.trace if t[0] == 'atp_rewards'][0][1:]
    if scale <= 64:
        assert np.abs(cands - e['cand_z'])[:, :, :3].max() < 1e-5
        assert np.abs(r0 - e['r0']).max() < 2e-4 and np.abs(r - e['r']).max() < 2e-5
        assert np.array_equal(lm_new, e['lm_after'])
        assert np.abs(z - e['z_after'])[:, :3].max() < 1e-5 and np.abs(xp - e['xp_after'])[:, :3].max() < 1e-5
        assert np.allclose(z[:, 3:], e['z_after'][:, 3:], rtol=1e-5, atol=1e-5 * scale)
    # the selection + re-batching + rescaling with the reference's rewards and candidates (:1203-1240)
    n_l = len(lm)
    big_z = np.concatenate([e['z_in']] + list(e['cand_z']))
    mixed = e['r0'].astype(np.float32) * np.float32(s / 250) + e['r'].astype(np.float32) * np.float32(250 - s / 250)
    order = np.argsort(-mixed, kind='stable')[:fx.B]
    picked = np.concatenate([big_z[i * (n_l // fx.B):(i + 1) * (n_l // fx.B)] for i in order])     # equally sized ligands
    assert np.allclose(picked[:, 3:] * 4.0, e['z_after'][:, 3:], rtol=1e-6, atol=1e-6 * scale)


def test_mixed_s30_event_matches_reference(fx, golden_weights):
    e = fx.ev['mixed30']
    orc = make_oracle(fx, golden_weights, fx.d0('mixed', 30))
    s_arr, t_arr = _arrs(fx, 30)
    lm = e['lm'].astype(np.int64)
    z, xp, lm_new = orc.mixed_event(30, s_arr, t_arr, e['z_in'], e['xp_in'], lm, fx.pm, 1e-3 * (30 / 500), 1e-3)
    cands = [t[1] for t in orc.trace if t[0] == 'cand']
    r0, r = [t for t in orc.trace if t[0] == 'atp_rewards'][0][1:]
    for i in range(4):
        assert np.abs(cands[i] - e[f'sub{i}_cand_z'])[:, :3].max() < 2e-5, i
        assert np.allclose(cands[i][:, 3:], e[f'sub{i}_cand_z'][:, 3:], rtol=1e-4, atol=1e-3), i
    # the chained structure: candidate i + 1 starts from candidate i's SPSA output, not from the entry state (:1286)
    assert np.abs(e['sub1_step_z_in'] - e['sub0_z_out']).max() == 0.0
    assert np.abs(e['sub0_step_z_in'] - e['z_in']).max() == 0.0
    assert [float(e[f'sub{i}_zeta']) for i in range(4)] == [6e-05, 6e-05, 1e-3, 1e-3]
    assert np.abs(r0 - e['r0']).max() < 5e-4 and np.abs(r - e['r']).max() < 5e-5
    assert np.array_equal(lm_new, e['lm_after'])
    assert np.abs(z - e['z_after'])[:, :3].max() < 2e-5 and np.abs(xp - e['xp_after'])[:, :3].max() < 2e-5


# ---- the inpainting loop's guidance branches (conditional_model.py:1570-1586 SPSA window, :1629-1778 ATP block) ----------
@pytest.fixture(scope='module')
def fxi():
    return Fixture('inpaint_b20')


def test_inpaint_event_schedule(fxi):
    kinds, s = fxi.index[:, 0], fxi.index[:, 1]
    assert sorted(s[kinds == 2].tolist()) == [12, 13, 14, 15, 16]              # SPSA: 12 <= s <= 16, first resampling only
    assert sorted(s[kinds == 1].tolist()) == [0, 2, 4, 6, 8, 10]               # ATP: s <= 10, s % 2 == 0
    assert (kinds == 0).sum() == 40                                            # 20 steps x 2 resamplings


def _arrs_T(fx, s, T):
    B = fx.B
    return (np.full((B, 1), s, np.float32) / np.float32(T), (np.full((B, 1), s, np.float32) + np.float32(1)) / np.float32(T))


def test_inpaint_spsa_event_matches_reference(fxi, golden_weights):
    e = fxi.ev['spsa16']
    assert abs(float(e['zeta']) - 1e-3 * (16 / 1200)) < 1e-12                  # zeta = 1e-3 * s / 1200 here (:1571-1572)
    orc = make_oracle(fxi, golden_weights, fxi.d0('spsa', 16))
    _, t_arr = _arrs_T(fxi, 16, 20)
    lm = e['lm'].astype(np.int64)
    z, xp = orc.my_update_z_lig(e['z_in'], e['xp_in'], lm, fxi.pm, t_arr, float(e['zeta']), float(e['guidance_scale']))
    fp, fm = [t for t in orc.trace if t[0] == 'spsa_rewards'][0][1:]
    assert np.abs(fp - e['f_plus']).max() < 2e-4 and np.abs(fm - e['f_minus']).max() < 2e-4
    assert np.abs(z - e['z_out'])[:, :3].max() < 2e-6 and np.abs(xp - e['xp_out'])[:, :3].max() < 2e-6


@pytest.mark.parametrize('s', [10])
def test_inpaint_atp_event_matches_reference(fxi, golden_weights, s):
    """The inpainting copy of the ATP block looks ahead from the current state with the ORIGINAL pocket (:1631)."""
    e = fxi.ev[f'atp{s}']
    orc = make_oracle(fxi, golden_weights, fxi.d0('atp', s))
    s_arr, t_arr = _arrs_T(fxi, s, 20)
    lm = e['lm'].astype(np.int64)
    assert np.abs(e['xp0'] - e['xp_in'])[:, :3].max() > 1e-3                   # the two pockets do differ
    z, xp, lm_new = orc.atp_event(s, s_arr, t_arr, e['z_in'], e['xp_in'], lm, fxi.pm, x0_pocket_group0=e['xp0'])
    r0, r = [t for t in orc.trace if t[0] == 'atp_rewards'][0][1:]
    cands = np.stack([t[1] for t in orc.trace if t[0] == 'cand'])
    scale = max(1.0, float(np.abs(e['z_in'][:, 3:]).max()))
    assert np.abs(cands - e['cand_z'])[:, :, :3].max() < 1e-4
    assert np.abs(r - e['r']).max() < 1e-4
    assert np.abs(r0 - e['r0']).max() < 5e-3 * max(1.0, np.log10(scale))       # look-ahead through a net fed 4^k-scaled features
    # with the wrong pocket for group 0 the look-ahead rewards of that group change
    orc2 = make_oracle(fxi, golden_weights, fxi.d0('atp', s))
    orc2.atp_event(s, s_arr, t_arr, e['z_in'], e['xp_in'], lm, fxi.pm)
    r0_wrong = [t for t in orc2.trace if t[0] == 'atp_rewards'][0][1]
    assert np.abs(r0_wrong[:fxi.B] - e['r0'][:fxi.B]).max() > 10 * np.abs(r0[:fxi.B] - e['r0'][:fxi.B]).max()
