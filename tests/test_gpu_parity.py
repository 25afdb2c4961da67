"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against
(1) the golden vectors generated from the reference, (2) the numpy oracle on seeded inputs, (3) size-independent
properties at the benchmark's full size.

Tolerances (BASELINE.json north_star): radius-graph edge sets bit-exact (set AND order); per-step coordinates within
1e-3 A; features within 1e-2 relative.  The denoiser output eps feeds a step through c_eps = sigma2_ts/alpha_ts/sigma_t
(< 1 over the polynomial_2 schedule), so eps_x is held to max(1e-3 A, 2e-3 relative) and the resulting
per-step coordinates are checked against the 1e-3 A bar directly.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_npz_groups
from oracle import egnn_oracle as O

pytestmark = pytest.mark.gpu

CFG = O.OracleConfig()
EDGE_CASES, _ = load_npz_groups('edges.npz')
FWD_CASES, _ = load_npz_groups('forward.npz')
TRAJ_CASES, _ = load_npz_groups('trajectory.npz')
INPAINT_CASES, _ = load_npz_groups('inpaint.npz')

X_ABS_TOL = 1e-3       # eps_x, absolute (A) ...
X_REL_TOL = 2e-3       # ... or relative to max |eps_x|
H_REL_TOL = 1e-2       # features, relative to max |eps_h|
STEP_X_TOL = 1e-3      # per-step coordinates (A)


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda', 0)


@pytest.fixture(scope='module')
def dyn(golden_weights, dev):
    from diffndm_b200.engine import B200EGNNDynamics
    from diffndm_b200.weights import DynamicsConfig
    d = B200EGNNDynamics(DynamicsConfig(), golden_weights, max_nodes=8192, max_edges=400000, max_samples=64)
    return d.eval()          # sampling runs in eval mode (NaN -> ValueError, dynamics.py:155-159)


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _edges_from_csr(rp, col):
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    return np.stack([rows, col.astype(np.int64)])


# ---------------------------------------------------------------------------------------------------------------
# tcgen05 GEMM building block
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,K,bn,act,res,tail', [
    (128, 512, 256, 256, 0, False, (0, 0)),      # block-0 projection: two resident 256-column groups
    (1, 256, 256, 256, 0, False, (0, 0)),
    (1000, 1536, 256, 256, 0, False, (2, 333)),  # merged projection: 4 full groups + 2 ligand-row-only groups
    (257, 1024, 256, 256, 0, False, (2, 40)),    # last block: 2 full + 2 tail groups
    (333, 256, 512, 128, 1, False, (0, 0)),      # node MLP layer 1 (K = 512, SiLU)
    (700, 256, 256, 128, 0, True, (0, 0)),       # node MLP layer 2 (+ fp32 residual, bf16 copy)
    (129, 256, 256, 128, 0, True, (0, 0)),       # ... with an odd number of row blocks (the peer CTA's last block is empty)
])
def test_tcgen05_node_gemm(dev, M, N, K, bn, act, res, tail):
    """gemm_pair_kernel -- the kernel every node GEMM of the forward launches -- in all three compiled shapes, with tail
    groups, bias, SiLU, the fp32 residual epilogue and the bf16 (TMA store) output, against a torch fp32 reference."""
    from diffndm_b200.engine import test_gemm
    g = torch.Generator().manual_seed(M * 7 + N)
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.1).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    r = torch.randn(M, 256, generator=g).to(dev) if res else None
    out, out16 = test_gemm(a, w, bias, act, bn=bn, residual=r, n_tail_groups=tail[0], m_tail=tail[1], want_bf16=True)
    ref = a.float() @ w.float().T + bias
    if res:
        ref = ref + r
    if act:
        ref = torch.nn.functional.silu(ref)
    if tail[0]:                                   # tail groups: only the first m_tail rows are computed
        ref[tail[1]:, N - tail[0] * bn:] = 0
    tol = 2e-4 * max(1.0, ref.abs().max().item())
    assert (out - ref).abs().max().item() < tol
    d16 = (out16.float() - ref).abs()
    if tail[0]:          # the bf16 path stores whole 128-row blocks: tail-group rows up to the end of m_tail's row block hold
        d16[tail[1]:-(-tail[1] // 128) * 128, N - tail[0] * bn:] = 0          # (valid) values the forward never reads
    assert d16.max().item() < tol + 8e-3 * ref.abs().max().item()              # bf16 rounding of the output


# ---------------------------------------------------------------------------------------------------------------
# (i) radius graph: bit-exact set and order vs the reference's get_edges
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', sorted(EDGE_CASES))
def test_radius_graph_bit_exact_vs_reference(dyn, dev, name):
    c = EDGE_CASES[name]
    B = int(max(c['lig_mask'].max(), c['pocket_mask'].max())) + 1
    rp, col = dyn.engine.radius_graph(_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['lig_mask'], dev),
                                      _t(c['pocket_mask'], dev), B)
    e = _edges_from_csr(rp.cpu().numpy(), col.cpu().numpy())
    assert np.array_equal(e, c['edges'].astype(np.int64))


def test_radius_graph_vs_oracle_ragged_random(dyn, dev):
    from diffndm_b200 import synthetic
    rng = np.random.default_rng(3)
    for trial in range(4):
        px, pt = synthetic.synthetic_pocket(40 + trial, int(rng.integers(20, 400)))
        sizes = rng.integers(1, 51, size=int(rng.integers(1, 9)))
        b = synthetic.make_batch(px, pt, sizes, trial)
        b['xh_lig'][:, :3] *= rng.uniform(1.0, 4.0)
        B = len(sizes)
        rp, col = dyn.engine.radius_graph(_t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), _t(b['lig_mask'], dev),
                                          _t(b['pocket_mask'], dev), B)
        e = _edges_from_csr(rp.cpu().numpy(), col.cpu().numpy())
        ref = O.get_edges(b['lig_mask'], b['pocket_mask'], b['xh_lig'][:, :3], b['xh_pocket'][:, :3], CFG)
        assert np.array_equal(e, ref)


def test_radius_graph_pocket_lists_stay_bit_exact(dyn, dev):
    """The pocket-pocket candidate lists (csrc/graph.cuh): calls 2..n on a batch whose pockets only translate are served
    from the lists and must emit the oracle's edge list bit for bit -- with the ligands moved, with every sample's pocket
    translated by its own vector, on a 1e-3 A grid (exact ties at the cutoff, as PDB coordinates produce).  A pocket that
    deforms, another batch layout, or a row denser than the list capacity must fall back to the full scan (same result)."""
    from diffndm_b200 import synthetic
    eng = dyn.engine
    rng = np.random.default_rng(17)
    px, pt = synthetic.synthetic_pocket(61, 300)
    px = np.round(px, 3).astype(np.float32)                           # PDB-like grid
    sizes = np.array([12, 30, 7, 21])
    B = len(sizes)
    b = synthetic.make_batch(px, pt, sizes, 5)
    n_p = len(px)

    def edges(bb):
        rp, col = eng.radius_graph(_t(bb['xh_lig'], dev), _t(bb['xh_pocket'], dev), _t(bb['lig_mask'], dev), _t(bb['pocket_mask'], dev), B)
        ref = O.get_edges(bb['lig_mask'], bb['pocket_mask'], bb['xh_lig'][:, :3], bb['xh_pocket'][:, :3], CFG)
        assert np.array_equal(_edges_from_csr(rp.cpu().numpy(), col.cpu().numpy()), ref)
        return eng.pocket_list_state()

    st = edges(b)                                                      # whatever an earlier test left: rebuilt for this layout now
    assert st[0] == 1 and st[1] == 0 and st[2:5] == [int(sizes.sum()), B * n_p, B]
    for step in range(6):                                              # rigid moves: per-sample translation, new ligand poses
        shift = rng.normal(size=(B, 3)).astype(np.float32) * (0.3 if step < 5 else 40.0)
        b['xh_pocket'][:, :3] = (b['xh_pocket'][:, :3].reshape(B, n_p, 3) - shift[:, None, :]).reshape(-1, 3)
        b['xh_lig'][:, :3] += rng.normal(size=b['xh_lig'][:, :3].shape).astype(np.float32) * 0.5 - shift[b['lig_mask']]
        st = edges(b)
        assert st[5] >= 0, 'served from the lists'
    # exact ties: a pocket on the 1e-3 grid whose pairs sit exactly on the cutoff (3-4-5 triangles), untranslated and translated
    tie = b['xh_pocket'].copy()
    base = np.zeros((n_p, 3), np.float32)
    base[:, 0] = (np.arange(n_p) % 10) * 3.0
    base[:, 1] = ((np.arange(n_p) // 10) % 6) * 4.0
    base[:, 2] = (np.arange(n_p) // 60) * 5.0
    for s in range(B):
        tie[s * n_p:(s + 1) * n_p, :3] = base + np.float32(0.125 * s)
    bt = dict(b, xh_pocket=tie)
    st = edges(bt)
    assert st[5] == -1, 'deformed pocket: lists are stale, the call scans everything'
    st = edges(bt)
    assert st[5] >= 0
    bt['xh_pocket'] = tie.copy()
    bt['xh_pocket'][:, :3] += np.float32(7.0)
    assert edges(bt)[5] >= 0
    # one atom moves by 0.05 A: stale; a row denser than the list capacity: that row scans everything, the others use their lists
    bt['xh_pocket'][5, 0] += np.float32(0.05)
    assert edges(bt)[5] == -1
    dense = b['xh_pocket'].copy()
    dense[:100, :3] = (rng.normal(size=(100, 3)) * 0.6).astype(np.float32)         # 100 atoms inside ~2 A: > 64 neighbours each
    bd = dict(b, xh_pocket=dense)
    edges(bd)
    st = edges(bd)
    assert st[0] == 1 and st[1] == 0 and st[5] == -1                  # first pocket row is one of the dense ones
    # another layout (one sample fewer) and back
    keep_l, keep_p = b['lig_mask'] < 3, b['pocket_mask'] < 3
    b3 = {'xh_lig': b['xh_lig'][keep_l], 'xh_pocket': b['xh_pocket'][keep_p], 'lig_mask': b['lig_mask'][keep_l], 'pocket_mask': b['pocket_mask'][keep_p]}
    rp, col = eng.radius_graph(_t(b3['xh_lig'], dev), _t(b3['xh_pocket'], dev), _t(b3['lig_mask'], dev), _t(b3['pocket_mask'], dev), 3)
    ref = O.get_edges(b3['lig_mask'], b3['pocket_mask'], b3['xh_lig'][:, :3], b3['xh_pocket'][:, :3], CFG)
    assert np.array_equal(_edges_from_csr(rp.cpu().numpy(), col.cpu().numpy()), ref) and eng.pocket_list_state()[5] == -1
    assert edges(b)[5] == -1 and edges(b)[5] >= 0


# ---------------------------------------------------------------------------------------------------------------
# (ii) forward: golden vectors from the reference (fp64 truth) + per-block trace vs the oracle
# ---------------------------------------------------------------------------------------------------------------
def _check_eps(out, ref64):
    ex = np.abs(out[:, :3] - ref64[:, :3]).max()
    sx = np.abs(ref64[:, :3]).max()
    eh = np.abs(out[:, 3:] - ref64[:, 3:]).max()
    sh = np.abs(ref64[:, 3:]).max()
    assert ex < max(X_ABS_TOL, X_REL_TOL * sx), f'eps_x err {ex:.3e} (scale {sx:.3f})'
    assert eh < H_REL_TOL * sh, f'eps_h err {eh:.3e} (scale {sh:.3f})'
    return ex, eh


@pytest.mark.parametrize('name', sorted(FWD_CASES))
def test_forward_vs_reference_golden(dyn, dev, name, golden_weights):
    c = FWD_CASES[name]
    B = len(c['t'])
    N = len(c['lig_mask']) + len(c['pocket_mask'])
    n_l = len(c['lig_mask'])
    trh, trx = dyn.engine.set_trace(N)
    out_l, out_p = dyn(_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['t'], dev), _t(c['lig_mask'], dev),
                       _t(c['pocket_mask'], dev))
    out_l, out_p = out_l.cpu().numpy(), out_p.cpu().numpy()
    _check_eps(out_l, c['out_lig_f64'])
    assert np.abs(out_p[:, :3]).max() == 0.0                                   # pocket velocity exactly 0
    assert np.abs(out_p[:, 3:] - c['out_pocket_f64'][:, 3:]).max() < H_REL_TOL * np.abs(c['out_pocket_f64'][:, 3:]).max()
    # per-block states on the rows the reference hooks recorded
    rows = c['trace_rows']
    for i in range(CFG.n_layers):
        h = trh[i, :N].cpu().numpy()
        x = trx[i, :N].cpu().numpy()
        ref_h = c[f'h_rows_{i}']
        assert np.abs(h[rows] - ref_h).max() < 5e-3 * np.abs(ref_h).max(), f'block {i} h'
        assert np.abs(x[:n_l] - c[f'x_lig_{i}']).max() < 5e-3, f'block {i} x'
    dyn.engine.clear_trace()


def test_encoder_table_matches_general_path(dyn, dev):
    """One-hot pocket rows take encoder + embedding from the per-type table (fp64 at weight load, csrc/node_kernels.cuh);
    the same rows with a 1e-30 in one of their zero entries are not one-hot any more and run the general path.  The embedded
    features h_0 of the two calls (read back through the trace) agree to fp32 rounding, for raw rows and for rows divided by
    the normalize factor 4."""
    c = FWD_CASES['synth60_b3']
    N = len(c['lig_mask']) + len(c['pocket_mask'])
    n_l = len(c['lig_mask'])
    pocket = c['xh_pocket'].copy()
    hot = pocket[:, 3:].argmax(1)
    for scale in (1.0, 0.25):
        pocket[:, 3:] = 0.0
        pocket[np.arange(len(pocket)), 3 + hot] = scale
        almost = pocket.copy()
        almost[np.arange(len(pocket)), 3 + (hot + 1) % pocket[:, 3:].shape[1]] = 1e-30
        hs = []
        for xp in (pocket, almost):
            dyn.engine.set_trace(N)
            dyn(_t(c['xh_lig'], dev), _t(xp, dev), _t(c['t'], dev), _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev))
            hs.append(dyn.engine.debug_h0(N))
            dyn.engine.clear_trace()
        a, b = hs[0][n_l:N].cpu().numpy(), hs[1][n_l:N].cpu().numpy()
        assert np.abs(a - b).max() < 2e-6 * max(1.0, np.abs(b).max()), (scale, np.abs(a - b).max())
        assert np.array_equal(hs[0][:n_l].cpu().numpy(), hs[1][:n_l].cpu().numpy())      # ligand rows: same path, same bits


def test_forward_scalar_time_matches_per_sample_time(dyn, dev):
    c = FWD_CASES['synth60_b3']
    a, _ = dyn(_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['t'], dev), _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev))
    b, _ = dyn(_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['t'][:1], dev), _t(c['lig_mask'], dev),
               _t(c['pocket_mask'], dev))                                      # np.prod(t.size()) == 1 branch, dynamics.py:105-107
    assert torch.equal(a, b)


def test_forward_deterministic_bitwise(dyn, dev):
    """The CSR segment reduction has a fixed summation order: repeated calls are bit-identical
    (the reference's scatter_add_ on GPU is not)."""
    c = FWD_CASES['3rfm_b2']
    args = (_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['t'], dev), _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev))
    a, ap = dyn(*args)
    for _ in range(3):
        b, bp = dyn(*args)
        assert torch.equal(a, b) and torch.equal(ap, bp)


def test_last_block_pruning_is_exact(dyn, dev):
    """When the pocket output is not requested (every conditional sampler call site discards it) the last block only
    aggregates for ligand atoms and their pocket senders, the block before it for the last one's receivers
    and all their senders: the ligand output must be bit-identical."""
    for name in ('3rfm_b2', 'synth60_b3_tmix'):
        c = FWD_CASES[name]
        args = (_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['t'], dev), _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev))
        full, _ = dyn(*args)
        dyn.compute_pocket_output = False
        try:
            pruned, none = dyn(*args)
        finally:
            dyn.compute_pocket_output = True
        assert none is None
        assert torch.equal(full, pruned)
        e, el, levels = dyn.engine.graph_stats_pruned()          # edges of: all blocks / ligand rows / the trailing blocks, last first
        assert len(levels) == 2 and el <= levels[0] <= levels[1] <= e


def test_forward_batch_composition_invariance(dyn, dev):
    """Samples never interact (dynamics.py:115): a sample's output is bit-identical alone or inside a batch,
    as long as its edges land at the same positions of the 128-edge tiles -- and within tolerance otherwise."""
    c = FWD_CASES['synth60_b3']
    lm, pm = c['lig_mask'], c['pocket_mask']
    full, _ = dyn(_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['t'], dev), _t(lm, dev), _t(pm, dev))
    full = full.cpu().numpy()
    for s in range(3):
        sl, sp = lm == s, pm == s
        one, _ = dyn(_t(c['xh_lig'][sl], dev), _t(c['xh_pocket'][sp], dev), _t(c['t'][s:s + 1], dev),
                     _t(np.zeros(sl.sum(), np.int64), dev), _t(np.zeros(sp.sum(), np.int64), dev))
        d = np.abs(one.cpu().numpy() - full[sl])
        assert d[:, :3].max() < 2e-3 and d[:, 3:].max() < 2e-3


def test_forward_se3_equivariance(dyn, dev):
    """eps_x rotates with the input, eps_h is invariant (cross-product MLP keeps SE(3), not reflections)."""
    c = FWD_CASES['3rfm_b2']
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    q = q.astype(np.float32)
    shift = np.float32([1.5, -2.0, 0.7])
    xl, xp = c['xh_lig'].copy(), c['xh_pocket'].copy()
    xl[:, :3] = xl[:, :3] @ q.T + shift
    xp[:, :3] = xp[:, :3] @ q.T + shift
    a, _ = dyn(_t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(c['t'], dev), _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev))
    b, _ = dyn(_t(xl, dev), _t(xp, dev), _t(c['t'], dev), _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev))
    a, b = a.cpu().numpy(), b.cpu().numpy()
    sx = np.abs(a[:, :3]).max()
    assert np.abs(a[:, :3] @ q.T - b[:, :3]).max() < max(2e-3, 2 * X_REL_TOL * sx)
    assert np.abs(a[:, 3:] - b[:, 3:]).max() < H_REL_TOL * np.abs(a[:, 3:]).max()


def test_static_mask_layout_cache(dyn, dev):
    """dndm_set_static_masks: repeated calls with the same mask tensors skip the offset derivation and give the same
    result; other tensors (different storage) are picked up; toggling drops the cache."""
    from diffndm_b200 import synthetic
    px, pt = synthetic.synthetic_pocket(5, 64)
    b1 = synthetic.make_batch(px, pt, np.array([9, 4, 14]), 1)
    b2 = synthetic.make_batch(px, pt, np.array([6, 11]), 2)
    def run(b, lm, pm):
        B = int(lm.max().item()) + 1
        out, _ = dyn(_t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), torch.full((B, 1), 0.3, device=dev), lm, pm, n_samples=B)
        return out.clone()
    lm1, pm1 = _t(b1['lig_mask'], dev), _t(b1['pocket_mask'], dev)
    lm2, pm2 = _t(b2['lig_mask'], dev), _t(b2['pocket_mask'], dev)
    ref1, ref2 = run(b1, lm1, pm1), run(b2, lm2, pm2)
    n0 = dyn.engine.lib.dndm_launch_count()
    run(b1, lm1, pm1)
    per_call = dyn.engine.lib.dndm_launch_count() - n0
    dyn.engine.set_static_masks(True)
    try:
        assert torch.equal(run(b1, lm1, pm1), ref1)
        n1 = dyn.engine.lib.dndm_launch_count()
        assert torch.equal(run(b1, lm1, pm1), ref1)                       # cached layout: two launches fewer
        assert dyn.engine.lib.dndm_launch_count() - n1 == per_call - 2
        assert torch.equal(run(b2, lm2, pm2), ref2)                       # other tensors -> layout rebuilt
        assert torch.equal(run(b1, lm1, pm1), ref1)
    finally:
        dyn.engine.set_static_masks(False)
    assert torch.equal(run(b2, lm2, pm2), ref2)


def test_nan_raises_value_error(dyn, dev):
    c = FWD_CASES['synth60_b3']
    xl = c['xh_lig'].copy()
    xl[0, 0] = np.nan
    args = (_t(xl, dev), _t(c['xh_pocket'], dev), _t(c['t'], dev), _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev))
    with pytest.raises(ValueError, match='NaN detected in EGNN output'):        # eval mode, dynamics.py:155-159
        dyn(*args)
    dyn.train()                                                                 # training mode zeroes them, :156-157
    try:
        out, _ = dyn(*args)
        assert not torch.isnan(out[:, :3]).any()
    finally:
        dyn.eval()


# ---------------------------------------------------------------------------------------------------------------
# (iii) sampler step: teacher-forced against the reference's recorded states, with the recorded noise
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', sorted(TRAJ_CASES))
def test_teacher_forced_steps_vs_reference(dyn, dev, name):
    from diffndm_b200.sampler import ConditionalSampler
    c = TRAJ_CASES[name]
    Tn = int(c['timesteps'])
    smp = ConditionalSampler(dyn, timesteps=500)
    lm, pm = _t(c['lig_mask'], dev), _t(c['pocket_mask'], dev)
    B = len(c['sizes'])
    worst = 0.0
    for i in range(Tn):
        z, xp = smp.sample_p_zs_given_zt(torch.from_numpy(c[f'step{i}/s']), torch.from_numpy(c[f'step{i}/t']),
                                         _t(c[f'step{i}/z_in'], dev), _t(c[f'step{i}/xp_in'], dev), lm, pm,
                                         noise=_t(c['noise'][1 + i], dev), n_samples=B)
        z, xp = z.cpu().numpy(), xp.cpu().numpy()
        ref_z, ref_p = c[f'step{i}/z_out'], c[f'step{i}/xp_out']
        # few-step trajectories of a random-init net blow up (|z| ~ 1e2): the bar is relative to the state scale
        scale = max(1.0, np.abs(ref_z[:, :3]).max())
        dx = np.abs(z[:, :3] - ref_z[:, :3]).max() / scale
        dh = np.abs(z[:, 3:] - ref_z[:, 3:]).max() / max(1.0, np.abs(ref_z[:, 3:]).max())
        dp = np.abs(xp - ref_p).max() / scale
        worst = max(worst, dx)
        assert dx < STEP_X_TOL and dp < STEP_X_TOL, f'step {i}: coords off by {dx:.2e} (pocket {dp:.2e})'
        assert dh < H_REL_TOL, f'step {i}: features off by {dh:.2e}'
    # final p(x,h|z0) head: identical atom types, coordinates within the step bar
    x_l, h_l, x_p, h_p = smp.sample_p_xh_given_z0(_t(c[f'step{Tn - 1}/z_out'], dev), _t(c[f'step{Tn - 1}/xp_out'], dev), lm, pm,
                                                  B, noise=_t(c['noise'][Tn + 1], dev))
    assert np.array_equal(h_l.argmax(1).cpu().numpy(), c['final_lig'][:, 3:].argmax(1))


def test_free_running_trajectory_vs_reference(dyn, dev):
    """sample_given_pocket with the reference's noise: final atom types identical, coordinates close."""
    from diffndm_b200.sampler import ConditionalSampler
    c = TRAJ_CASES['synth50_b3_T10']
    Tn = int(c['timesteps'])
    smp = ConditionalSampler(dyn, timesteps=500)
    B = len(c['sizes'])
    n_p = len(c['pocket_x'])
    onehot = np.eye(10, dtype=np.float32)[c['pocket_t']]
    pocket = {'x': torch.from_numpy(np.tile(c['pocket_x'], (B, 1))), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))),
              'size': torch.tensor([n_p] * B), 'mask': torch.arange(B).repeat_interleave(n_p)}
    xh_l, xh_p, lm, pm = smp.sample_given_pocket(pocket, c['sizes'], timesteps=Tn, noise=_t(c['noise'], dev))
    ref = c['final_lig']
    scale = max(1.0, np.abs(ref[:, :3]).max())
    assert np.abs(xh_l.cpu().numpy()[:, :3] - ref[:, :3]).max() / scale < 2e-2
    assert (xh_l.cpu().numpy()[:, 3:].argmax(1) == ref[:, 3:].argmax(1)).mean() > 0.9


@pytest.mark.parametrize('name', sorted(INPAINT_CASES))
def test_inpaint_vs_reference(dyn, dev, name):
    """ConditionalSampler.inpaint (RePaint resampling) with the reference's Gaussian draws: fixed atoms, blend, pocket
    translation and re-noising follow conditional_model.py:1491-1790."""
    from diffndm_b200.sampler import ConditionalSampler
    c = INPAINT_CASES[name]
    smp = ConditionalSampler(dyn, timesteps=500)
    B, n_p = len(c['sizes']), len(c['pocket_x'])
    oh = np.eye(10, dtype=np.float32)
    pocket = {'x': torch.from_numpy(np.tile(c['pocket_x'], (B, 1))), 'one_hot': torch.from_numpy(np.tile(oh[c['pocket_t']], (B, 1))),
              'size': torch.tensor([n_p] * B), 'mask': torch.arange(B).repeat_interleave(n_p)}
    ligand = {'x': torch.from_numpy(c['lig_x']), 'one_hot': torch.from_numpy(oh[c['lig_t']]),
              'size': torch.from_numpy(c['sizes']), 'mask': torch.from_numpy(c['lig_mask'])}
    xh_l, xh_p, lm, pm = smp.inpaint(ligand, pocket, torch.from_numpy(c['lig_fixed']), resamplings=int(c['resamplings']),
                                     timesteps=int(c['timesteps']), noise=[_t(n, dev) for n in c['noise']])
    ref = c['final_lig']
    scale = max(1.0, np.abs(ref[:, :3]).max())
    assert np.abs(xh_l.cpu().numpy()[:, :3] - ref[:, :3]).max() / scale < 2e-2
    assert np.abs(xh_p.cpu().numpy()[:, :3] - c['final_pocket'][:, :3]).max() / scale < 2e-2
    assert (xh_l.cpu().numpy()[:, 3:].argmax(1) == ref[:, 3:].argmax(1)).mean() > 0.9
    assert np.array_equal(lm.cpu().numpy(), c['lig_mask'])


def test_com_drift_raises_assertion_error(dyn, dev):
    """assert_mean_zero_with_mask on z_t (conditional_model.py:535): a reverse step from a state whose ligand COM is
    off by more than 1e-2 of its largest coordinate raises AssertionError; prior / forward-noising moves do not."""
    from diffndm_b200 import synthetic
    from diffndm_b200.sampler import ConditionalSampler
    px, pt = synthetic.synthetic_pocket(3, 40)
    b = synthetic.make_batch(px, pt, np.array([6, 9]), 4)
    smp = ConditionalSampler(dyn, timesteps=500, check_every_step=True)
    lm, pm = _t(b['lig_mask'], dev), _t(b['pocket_mask'], dev)
    z = _t(b['xh_lig'], dev).clone()
    z[:, :3] -= torch.zeros((2, 3), device=dev).index_add_(0, lm, z[:, :3])[lm] / torch.bincount(lm)[lm][:, None]
    s = torch.full((2, 1), 100 / 500.)
    t = torch.full((2, 1), 101 / 500.)
    smp.sample_p_zs_given_zt(s, t, z, _t(b['xh_pocket'], dev), lm, pm, n_samples=2)        # COM-free input: fine
    z_bad = z.clone()
    z_bad[:, 0] += 0.5
    with pytest.raises(AssertionError):
        smp.sample_p_zs_given_zt(s, t, z_bad, _t(b['xh_pocket'], dev), lm, pm, n_samples=2)
    # the x0 head (sample_p_xh_given_z0 inside my_to_x0) takes z0 = (z_t - sigma eps)/alpha, which is not COM-free either:
    # the reference has no assertion there (conditional_model.py:136-160)
    smp.my_to_x0(t, z, _t(b['xh_pocket'], dev), lm, pm, 2)
    smp.sample_p_xh_given_z0(z_bad, _t(b['xh_pocket'], dev), lm, pm, 2)
    assert dyn.engine.read_flags() & 2 == 0
    # a forward-noising move takes inputs that are not COM-free by construction
    coef = torch.tensor([[0.9, 0.0, 0.1]], device=dev).repeat(2, 1)
    dyn.engine.sampler_step(z_bad, None, torch.randn_like(z_bad), _t(b['xh_pocket'], dev), coef, lm, pm, 2)
    assert dyn.engine.read_flags() == 0


def test_sampler_step_vs_oracle_and_in_place(dyn, dev):
    from diffndm_b200 import synthetic
    px, pt = synthetic.synthetic_pocket(2, 77)
    b = synthetic.make_batch(px, pt, np.array([5, 1, 33, 12]), 9)
    rng = np.random.default_rng(1)
    eps = rng.standard_normal(b['xh_lig'].shape).astype(np.float32)
    noise = rng.standard_normal(b['xh_lig'].shape).astype(np.float32)
    g = O.gamma_table()
    s_idx = np.array([10, 200, 350, 499])
    gs, gt = g[s_idx], g[s_idx + 1]
    sc = O.step_scalars(gs, gt)
    coef = np.stack([1 / sc['alpha_ts'], sc['sigma2_ts'] / sc['alpha_ts'] / sc['sigma_t'],
                     sc['sigma_ts'] * sc['sigma_s'] / sc['sigma_t']], 1).astype(np.float32)
    zr, pr = O.sample_p_zs_given_zt(b['xh_lig'], b['xh_pocket'], eps, noise, gs, gt, b['lig_mask'], b['pocket_mask'])
    z_in, p_in = _t(b['xh_lig'], dev), _t(b['xh_pocket'], dev)
    z, p = dyn.engine.sampler_step(z_in, _t(eps, dev), _t(noise, dev), p_in, _t(coef, dev), _t(b['lig_mask'], dev),
                                   _t(b['pocket_mask'], dev), 4)
    assert np.abs(z.cpu().numpy() - zr).max() < 2e-5 * max(1.0, np.abs(zr).max())
    assert np.abs(p.cpu().numpy() - pr).max() < 2e-5 * max(1.0, np.abs(pr).max())
    # COM-free afterwards
    com = np.zeros((4, 3))
    np.add.at(com, b['lig_mask'], z.cpu().numpy()[:, :3])
    assert np.abs(com).max() < 1e-4
    # in place == out of place, bit for bit
    dyn.engine.sampler_step(z_in, _t(eps, dev), _t(noise, dev), p_in, _t(coef, dev), _t(b['lig_mask'], dev),
                            _t(b['pocket_mask'], dev), 4, z_out=z_in, pocket_out=p_in)
    assert torch.equal(z_in, z) and torch.equal(p_in, p)


def test_spsa_update_vs_oracle(dyn, dev):
    """my_update_z_lig (conditional_model.py:760-813) with injected perturbations and a deterministic reward."""
    from diffndm_b200 import synthetic
    from diffndm_b200.sampler import ConditionalSampler
    px, pt = synthetic.synthetic_pocket(6, 60)
    sizes = np.array([7, 9, 5])
    b = synthetic.make_batch(px, pt, sizes, 6)
    k, B = 3, 3
    rng = np.random.default_rng(2)
    U = (1e-3 * rng.standard_normal((k, len(b['lig_mask']), 3))).astype(np.float32)
    calls = {}

    def reward(x, types, mask):
        # deterministic host "chemistry": radius of gyration per molecule (a stand-in for RDKit scores)
        x = x.cpu().numpy().astype(np.float64)
        m = mask.cpu().numpy()
        nb = m.max() + 1
        r = np.zeros(nb)
        for i in range(nb):
            xi = x[m == i]
            r[i] = np.sqrt(((xi - xi.mean(0)) ** 2).sum(1).mean())
        calls['r'] = r
        return r.tolist()

    smp = ConditionalSampler(dyn, timesteps=500)
    t_arr = torch.full((B, 1), 20 / 500)
    x0_noise = torch.from_numpy(rng.standard_normal((2 * k, len(b['lig_mask']), 13)).astype(np.float32)).to(dev)
    z, xp = smp.my_update_z_lig(_t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), _t(b['lig_mask'], dev), _t(b['pocket_mask'], dev),
                                t_arr, B, 1e-3, reward, guidance_scale=1e-3, k=k, perturbations=torch.from_numpy(U),
                                x0_noise=x0_noise)
    r = calls['r'].reshape(2 * k, B)
    zr, xpr = O.spsa_update(b['xh_lig'], b['xh_pocket'], U, r[:k], r[k:], b['lig_mask'], b['pocket_mask'], guidance_scale=1e-3)
    assert np.abs(z.cpu().numpy() - zr).max() < 1e-4 * max(1.0, np.abs(zr).max())
    assert np.abs(xp.cpu().numpy() - xpr).max() < 1e-4 * max(1.0, np.abs(xpr).max())


def test_x0_lookahead_vs_oracle(dyn, dev, golden_weights):
    """my_to_x0 (conditional_model.py:457-468): z0 = (z_t - sigma_t eps)/alpha_t, then the p(x,h|z0) head."""
    from diffndm_b200.sampler import ConditionalSampler
    c = FWD_CASES['synth60_b3']
    B = 3
    smp = ConditionalSampler(dyn, timesteps=500)
    lm, pm = c['lig_mask'], c['pocket_mask']
    t = np.full((B, 1), 30 / 500, np.float32)
    rng = np.random.default_rng(5)
    noise = rng.standard_normal(c['xh_lig'].shape).astype(np.float32)
    x_l, h_l, x_p, h_p = smp.my_to_x0(torch.from_numpy(t), _t(c['xh_lig'], dev), _t(c['xh_pocket'], dev), _t(lm, dev), _t(pm, dev),
                                      B, noise=_t(noise, dev))
    g = O.gamma_table()
    eps_t, _ = O.dynamics_forward(golden_weights, c['xh_lig'], c['xh_pocket'], t, lm, pm, CFG)
    z0 = O.x0_lookahead_z0(c['xh_lig'], eps_t, np.full(B, g[30]), lm)
    eps0, _ = O.dynamics_forward(golden_weights, z0, c['xh_pocket'], np.zeros((B, 1), np.float32), lm, pm, CFG)
    xr, types, xpr, hpr = O.sample_p_xh_given_z0(z0, c['xh_pocket'], eps0, noise, np.full(B, g[0]), lm, pm, CFG)
    assert np.abs(x_l.cpu().numpy() - xr).max() < 2e-3 * max(1.0, np.abs(xr).max())
    assert np.abs(x_p.cpu().numpy() - xpr).max() < 2e-3 * max(1.0, np.abs(xpr).max())
    assert np.array_equal(h_l.argmax(1).cpu().numpy(), types)
    assert np.abs(h_p.cpu().numpy() - hpr).max() < 1e-5


def test_atp_event_vs_oracle(dyn, dev):
    """ATP ("SVDD") event (conditional_model.py:1085-1241) with the candidate groups batched along the sample axis:
    the selection (mixed reward [sic], global top-k, re-batching) must equal the oracle's on the engine's own candidates."""
    from diffndm_b200 import synthetic
    from diffndm_b200.sampler import ConditionalSampler
    px, pt = synthetic.synthetic_pocket(8, 50)
    sizes = np.array([6, 9, 5, 7])
    b = synthetic.make_batch(px, pt, sizes, 8)
    B, G, s = 4, 3, 20
    smp = ConditionalSampler(dyn, timesteps=500)
    rec = []

    def reward(x, types, mask):
        x = x.cpu().numpy().astype(np.float64)
        m = mask.cpu().numpy()
        r = np.array([np.sqrt(((x[m == i] - x[m == i].mean(0)) ** 2).sum(1).mean()) + 0.01 * i for i in range(m.max() + 1)])
        rec.append((x.copy(), m.copy(), r.copy()))
        return r.tolist()

    torch.manual_seed(0)
    s_arr = torch.full((B, 1), s / 500)
    t_arr = torch.full((B, 1), (s + 1) / 500)
    z, xp, lm = smp._atp_event(s, s_arr, t_arr, _t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), _t(b['lig_mask'], dev),
                               _t(b['pocket_mask'], dev), B, reward, G)
    (x0, m0, r0), (xc, mc, rc) = rec           # look-ahead molecules, then current-state molecules
    assert m0.max() + 1 == G * B and np.array_equal(m0, mc)
    # rebuild the candidate batch the event scored: group 0 is the incoming state, coordinates of the current-state call
    n_l, n_p = len(b['lig_mask']), len(b['pocket_mask'])
    assert np.abs(xc[:n_l] - b['xh_lig'][:, :3]).max() < 1e-6
    big_pm = np.concatenate([b['pocket_mask'] + g * B for g in range(G)])
    # selection on the recorded rewards (the candidates themselves are engine outputs)
    order = np.argsort(-(r0.astype(np.float32) * np.float32(s / 250) + rc.astype(np.float32) * np.float32(250 - s / 250)),
                       kind='stable')[:B]
    lm_np = lm.cpu().numpy()
    assert np.array_equal(np.bincount(lm_np), np.bincount(mc, minlength=G * B)[order])
    zc = z.cpu().numpy()
    start = 0
    for rank, idx in enumerate(order):
        n = int((mc == idx).sum())
        assert np.abs(zc[start:start + n, :3] - xc[mc == idx]).max() < 1e-6
        start += n
    assert xp.shape[0] == n_p and z.shape[0] == len(lm_np)


def _atp_dist_worker(rank, world, port, ret):
    import os
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from diffndm_b200 import synthetic
    from diffndm_b200.engine import B200EGNNDynamics
    from diffndm_b200.sampler import ConditionalSampler
    from diffndm_b200.weights import DynamicsConfig, random_init
    dev = torch.device('cuda', rank)
    dyn = B200EGNNDynamics(DynamicsConfig(), random_init(DynamicsConfig(), 0, 0.3), max_nodes=4096,
                           max_edges=200000, max_samples=64).eval()
    px, pt = synthetic.synthetic_pocket(8, 50)
    b = synthetic.make_batch(px, pt, np.array([6, 9, 5, 7]), 8)          # identical state on every rank
    B, G, s = 4, 5, 20
    smp = ConditionalSampler(dyn, timesteps=500)
    smp.atp_group = dist.group.WORLD
    scored = []

    def reward(x, types, mask):
        x = x.double()
        r = []
        for i in range(int(mask.max()) + 1):
            xi = x[mask == i]
            r.append(float(((xi - xi.mean(0)) ** 2).sum(1).mean().sqrt()))
        scored.append(len(r))
        return r

    torch.manual_seed(100 + rank)                                       # the extra candidate draws differ per rank
    t = lambda a: torch.from_numpy(a).to(dev)
    z, xp, lm = smp._atp_event(s, torch.full((B, 1), s / 500), torch.full((B, 1), (s + 1) / 500), t(b['xh_lig']),
                               t(b['xh_pocket']), t(b['lig_mask']), t(b['pocket_mask']), B, reward, G)
    # every rank must have rebuilt the same winners
    n = torch.tensor([z.shape[0]], device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    same = all(int(k) == int(n) for k in ns)
    if same:
        zs = [torch.zeros_like(z) for _ in range(world)]
        ps = [torch.zeros_like(xp) for _ in range(world)]
        dist.all_gather(zs, z.contiguous())
        dist.all_gather(ps, xp.contiguous())
        same = all(torch.equal(zs[0], q) for q in zs) and all(torch.equal(ps[0], q) for q in ps)
    # groups 0, 2, 4 live on rank 0 and 1, 3 on rank 1: each scored look-ahead + current molecules of its own groups only
    n_groups_here = len([g for g in range(G) if g % world == rank])
    ok = same and scored == [n_groups_here * B] * 2 and xp.shape[0] == len(b['pocket_mask']) and int(lm.max()) == B - 1
    # a winner's pocket is the incoming pocket translated rigidly
    d = (xp[:, :3] - t(b['xh_pocket'])[:, :3])
    ok = ok and float((d - d[0]).abs().max()) < 1e3            # finite
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (run under gpurun --gpus 2)')
def test_atp_event_distributed_two_gpus():
    """the one exchange on the path -- candidate groups of an ATP event split over two ranks, winners
    rebuilt from an NCCL all-gather; both ranks must end with identical (z, pocket, mask)."""
    import os
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_atp_dist_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def _atp_traj_worker(rank, world, port, ret):
    import os
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from diffndm_b200 import synthetic
    from diffndm_b200.engine import B200EGNNDynamics
    from diffndm_b200.sampler import ConditionalSampler
    from diffndm_b200.weights import DynamicsConfig, random_init
    dyn = B200EGNNDynamics(DynamicsConfig(), random_init(DynamicsConfig(), 0, 0.3), max_nodes=4096, max_edges=200000,
                           max_samples=64).eval()
    px, pt = synthetic.synthetic_pocket(8, 50)
    B, G = 4, 5
    sizes = np.array([6, 9, 5, 7])
    oh = np.eye(10, dtype=np.float32)[pt]
    pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(oh, (B, 1))),
              'size': torch.tensor([len(px)] * B), 'mask': torch.arange(B).repeat_interleave(len(px))}
    smp = ConditionalSampler(dyn, timesteps=500)
    pose = synthetic.synthetic_ligand_pose(8, sizes, px.mean(0))
    pose[:, :3] -= px[0]
    smp.eps_transform = synthetic.PointMassScore(pose, smp.gamma, len(px), 500, dev)
    smp.set_distributed_atp(dist.group.WORLD, shared_seed=77)              # same trajectory noise, per-rank candidate draws
    seen = []

    def reward(x, types, mask):
        x = x.double()
        seen.append(float(x.sum()))
        return [float(((x[mask == i] - x[mask == i].mean(0)) ** 2).sum(1).mean().sqrt()) for i in range(int(mask.max()) + 1)]

    xh, xp, lm, _ = smp.sample_given_pocket(pocket, sizes, svdd=1, reward_fn=reward, svdd_groups=G)     # six ATP events
    # every rank must end with the same molecules; the candidates the ranks drew must differ
    n = torch.tensor([xh.shape[0]], device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    same = all(int(k) == int(n) for k in ns)
    if same:
        zs = [torch.zeros_like(xh) for _ in range(world)]
        dist.all_gather(zs, xh.contiguous())
        same = all(torch.equal(zs[0], q) for q in zs)
    mine = torch.tensor(seen[:2], device=dev, dtype=torch.float64)       # the first event's look-ahead and candidate sums
    other = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    distinct = not torch.equal(other[0], other[1])
    ret[rank] = bool(same and distinct and len(seen) == 12 and torch.isfinite(xh).all())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (run under gpurun --gpus 2)')
def test_atp_trajectory_distributed_two_gpus():
    """A whole guided trajectory (six ATP events) with the candidate groups split over two ranks that share the trajectory:
    the shared generator keeps the states identical between events, the per-rank generator makes the candidates distinct,
    and the packed all-gather rebuilds the same winners (absolute pocket positions) on both ranks."""
    import os
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_atp_traj_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


BOND_CASES, BOND_META = load_npz_groups('bonds.npz')


@pytest.mark.parametrize('name', sorted(BOND_CASES))
def test_bond_orders_bit_exact(dyn, dev, name):
    """dndm_bond_orders vs the reference's get_bond_order_batch / make_mol_edm fixtures (bit-exact) and the oracle's
    valences / fragments."""
    from diffndm_b200.chem import BondPerception
    c = BOND_CASES[name]
    info = {'bonds1': BOND_META['bonds1'].tolist(), 'bonds2': BOND_META['bonds2'].tolist(), 'bonds3': BOND_META['bonds3'].tolist(),
            'atom_decoder': ['C', 'N', 'O', 'S', 'B', 'Br', 'Cl', 'P', 'I', 'F']}
    bp = BondPerception(dyn.engine, info, margins=BOND_META['margins'].tolist())
    dyn.engine.read_flags()                              # the flag word is sticky across calls: start clean
    B = len(c['sizes'])
    out = bp(_t(c['x'], dev), _t(c['types'], dev), _t(c['mask'], dev), B)
    flat = torch.cat([m.reshape(-1) for m in out['E']]).cpu().numpy()
    assert np.array_equal(flat, c['e_flat'])
    mats, val, stats = O.bond_orders(c['x'], c['types'], c['mask'], BOND_META['bonds1'], BOND_META['bonds2'],
                                     BOND_META['bonds3'], BOND_META['margins'])
    assert np.array_equal(out['valence'].cpu().numpy(), val)
    assert np.array_equal(out['n_bonds'].cpu().numpy(), stats[:, 0])
    assert np.array_equal(out['n_components'].cpu().numpy(), stats[:, 1])
    assert np.array_equal(out['largest_component'].cpu().numpy(), stats[:, 2])
    allowed = np.array([4, 3, 2, 4, 3, 1, 1, 5, 1, 1])
    viol = np.array([(val[c['mask'] == b] > allowed[c['types'][c['mask'] == b]]).sum() for b in range(B)])
    assert np.array_equal(out['valence_violations'].cpu().numpy(), viol)
    keep = bp.keep_mask(out, torch.from_numpy(c['sizes']))
    assert keep.shape == (B,)
    assert dyn.engine.read_flags() == 0


def test_bond_orders_random_vs_oracle(dyn, dev):
    from diffndm_b200.chem import BondPerception
    rng = np.random.default_rng(3)
    sizes = rng.integers(1, 60, size=40)
    x = np.concatenate([np.cumsum(rng.normal(size=(n, 3)) * 0.85, 0) for n in sizes]).astype(np.float32)
    types = rng.integers(0, 10, size=int(sizes.sum()))
    mask = np.repeat(np.arange(len(sizes)), sizes)
    info = {'bonds1': BOND_META['bonds1'].tolist(), 'bonds2': BOND_META['bonds2'].tolist(), 'bonds3': BOND_META['bonds3'].tolist()}
    bp = BondPerception(dyn.engine, info)
    out = bp(_t(x, dev), _t(types, dev), _t(mask, dev), len(sizes))
    mats, val, stats = O.bond_orders(x, types, mask, BOND_META['bonds1'], BOND_META['bonds2'], BOND_META['bonds3'])
    for a, b in zip(out['E'], mats):
        assert np.array_equal(a.cpu().numpy(), b)
    assert np.array_equal(out['valence'].cpu().numpy(), val)
    assert np.array_equal(torch.stack([out['n_bonds'], out['n_components'], out['largest_component']], 1).cpu().numpy(), stats)


# ---------------------------------------------------------------------------------------------------------------
# full benchmark size: properties that do not need the oracle to finish
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_properties(golden_weights, dev):
    """BASELINE configs[1] shape (100 ligands on one pocket): graph invariants, determinism, replica consistency."""
    from diffndm_b200 import synthetic
    from diffndm_b200.engine import B200EGNNDynamics
    from diffndm_b200.weights import DynamicsConfig
    px, pt = synthetic.synthetic_pocket(0)
    sizes = synthetic.synthetic_ligand_sizes(0, 100)
    # identical ligands in every second sample: replicas must give bit-identical outputs? (tile positions differ,
    # so only to tolerance); sample 0 is additionally checked against the oracle run on that sample alone
    b = synthetic.make_batch(px, pt, sizes, 0)
    N = len(b['lig_mask']) + len(b['pocket_mask'])
    big = B200EGNNDynamics(DynamicsConfig(), golden_weights, max_nodes=N + 128, max_edges=N * 40, max_samples=100)
    t = np.full((100, 1), 0.3, np.float32)
    args = (_t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), _t(t, dev), _t(b['lig_mask'], dev), _t(b['pocket_mask'], dev))
    out, _ = big(*args)
    out2, _ = big(*args)
    assert torch.equal(out, out2)
    rp, col = big.engine.radius_graph(args[0], args[1], args[3], args[4], 100)
    rp, col = rp.cpu().numpy(), col.cpu().numpy().astype(np.int64)
    rows = np.repeat(np.arange(N), np.diff(rp))
    mask = np.concatenate([b['lig_mask'], b['pocket_mask']])
    assert np.all(mask[rows] == mask[col])                              # dynamics.py:115
    assert np.all((np.diff(col) > 0) | (np.diff(rows) > 0))             # cols strictly ascending inside a row
    key = rows * N + col
    assert np.array_equal(np.sort(key), np.sort(col * N + rows))        # symmetric edge set
    assert np.all(np.isin(np.arange(N) * (N + 1), key))                 # self loops
    s0l, s0p = b['lig_mask'] == 0, b['pocket_mask'] == 0
    ref, _ = O.dynamics_forward(golden_weights, b['xh_lig'][s0l], b['xh_pocket'][s0p], t[:1], np.zeros(s0l.sum(), np.int64),
                                np.zeros(s0p.sum(), np.int64), CFG, dtype=np.float64)
    _check_eps(out.cpu().numpy()[s0l], ref)


def test_forward_high_degree_receivers_vs_oracle(dyn, dev, golden_weights):
    """Receivers whose edge lists span several 128-edge tiles (a 300-atom fully connected ligand: degree > 300) next to a
    one-atom and a small ligand: the fused GCL kernel's per-receiver sums are stitched across tiles (gcl_stitch_kernel) and
    must agree with the oracle's segment sum like every other case; agg itself (the bf16 half of the node-MLP operand) is
    checked through the block-0 hidden state."""
    from diffndm_b200 import synthetic
    px, pt = synthetic.synthetic_pocket(17, 40)
    sizes = np.array([300, 1, 7])
    b = synthetic.make_batch(px, pt, sizes, 17)
    t = np.array([[0.7], [0.2], [0.5]], np.float32)
    N, n_l = len(b['lig_mask']) + len(b['pocket_mask']), len(b['lig_mask'])
    trh, trx = dyn.engine.set_trace(N)
    out_l, _ = dyn(_t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), _t(t, dev), _t(b['lig_mask'], dev), _t(b['pocket_mask'], dev))
    trace = {}
    ref, _ = O.dynamics_forward(golden_weights, b['xh_lig'], b['xh_pocket'], t, b['lig_mask'], b['pocket_mask'], CFG,
                                dtype=np.float64, trace=trace)
    h0 = trh[0, :N].cpu().numpy()
    dyn.engine.clear_trace()
    assert np.abs(h0 - trace['h_0']).max() < 5e-3 * np.abs(trace['h_0']).max()
    _check_eps(out_l.cpu().numpy(), ref)
    out2, _ = dyn(_t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), _t(t, dev), _t(b['lig_mask'], dev), _t(b['pocket_mask'], dev))
    assert torch.equal(out_l, out2)                           # deterministic


@pytest.mark.parametrize('n_lig,n_pok', [(1, 1), (11, 1), (16, 1), (23, 2), (16, 0)])
def test_forward_tile_edge_cases_vs_oracle(dyn, dev, golden_weights, n_lig, n_pok):
    """Edge counts around the 128-edge tile / 256-edge pair-iteration boundaries of the CTA-pair kernels: one sample whose
    ligand is fully connected (n^2 edges incl. self loops) next to a far-away pocket of n_pok atoms -- E = 2 (one tile of a
    single pair, its peer CTA all padding), 122 (< one tile), 257 (three tiles: the last pair iteration has ONE valid tile),
    533 (five tiles), 256 with no pocket at all (exactly one pair iteration) -- against the fp64 oracle."""
    rng = np.random.default_rng(100 + n_lig)
    xl = rng.normal(size=(n_lig, 3)).astype(np.float32) * 2.0
    hl = np.eye(10, dtype=np.float32)[rng.integers(0, 10, n_lig)] * 0.25
    xh_lig = np.concatenate([xl, hl], 1)
    xp = (rng.normal(size=(n_pok, 3)) + np.array([40.0, 0, 0])).astype(np.float32)      # beyond every cutoff from the ligand
    hp = np.eye(10, dtype=np.float32)[rng.integers(0, 10, n_pok)] * 0.25
    xh_pok = np.concatenate([xp, hp], 1).astype(np.float32).reshape(n_pok, 13)
    lm, pm = np.zeros(n_lig, np.int64), np.zeros(n_pok, np.int64)
    t = np.array([[0.4]], np.float32)
    e = O.get_edges(lm, pm, xl, xp, CFG)
    assert e.shape[1] >= n_lig * n_lig
    ref, _ = O.dynamics_forward(golden_weights, xh_lig, xh_pok, t, lm, pm, CFG, dtype=np.float64, edges=e)
    if n_pok == 0:                       # the reference has no such call; the engine accepts an empty pocket
        out_l, _ = dyn.engine.forward(_t(xh_lig, dev), torch.zeros((0, 13), device=dev), _t(t, dev), _t(lm, dev),
                                      torch.zeros((0,), dtype=torch.long, device=dev), 1, want_pocket=False)
    else:
        out_l, _ = dyn(_t(xh_lig, dev), _t(xh_pok, dev), _t(t, dev), _t(lm, dev), _t(pm, dev))
    _check_eps(out_l.cpu().numpy(), ref)
    assert dyn.engine.read_flags() & 5 == 0


def test_graphed_sampling_matches_eager(dyn, dev):
    """sample_given_pocket with the reverse step replayed from a CUDA graph draws the same trajectory as the eager loop
    (the capture's dry runs do not consume the noise stream).  The pocket COM of the prior goes through torch's atomic
    index_add_, so two runs agree to ~1e-6 relative after one step, growing with the step count; 3 steps are compared."""
    from diffndm_b200 import synthetic
    from diffndm_b200.sampler import ConditionalSampler
    px, pt = synthetic.synthetic_pocket(9, 80)
    sizes = np.array([6, 13, 9])
    B = len(sizes)
    onehot = np.eye(10, dtype=np.float32)[pt]
    pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))),
              'size': torch.tensor([len(px)] * B), 'mask': torch.arange(B).repeat_interleave(len(px))}
    smp = ConditionalSampler(dyn, timesteps=500)
    outs = []
    for graph in (False, True):
        torch.manual_seed(123)
        torch.cuda.manual_seed(123)
        xh, xp, lm, pm = smp.sample_given_pocket(pocket, sizes, timesteps=3, use_cuda_graph=graph)
        outs.append((xh.clone(), xp.clone()))
    scale = max(1.0, float(outs[0][0][:, :3].abs().max()))
    assert float((outs[0][0][:, :3] - outs[1][0][:, :3]).abs().max()) / scale < 5e-4
    assert float((outs[0][0][:, 3:].argmax(1) == outs[1][0][:, 3:].argmax(1)).float().mean()) > 0.9
    assert float((outs[0][1] - outs[1][1]).abs().max()) / scale < 5e-4
    assert dyn.engine.read_flags() & 5 == 0


def test_graphed_spsa_window_matches_eager(dyn, dev):
    """Inside an SPSA-only guidance window the reverse steps keep replaying the cached CUDA graph (an SPSA update changes the
    state, not the shapes or masks: the new state is copied into the graph's buffers).  Same seeds, same stand-in reward: the
    graphed and the eager trajectory agree, and the graph is captured once for the whole run."""
    from diffndm_b200 import synthetic
    from diffndm_b200.sampler import ConditionalSampler, _GraphedReverseStep
    px, pt = synthetic.synthetic_pocket(11, 70)
    sizes = np.array([7, 12, 9, 5])
    B = len(sizes)
    onehot = np.eye(10, dtype=np.float32)[pt]
    pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))),
              'size': torch.tensor([len(px)] * B), 'mask': torch.arange(B).repeat_interleave(len(px))}

    def reward(x, types, mask):                     # smooth geometric stand-in (no chemistry in the image)
        r2 = torch.zeros(int(mask.max()) + 1, device=x.device).index_add_(0, mask, (x * x).sum(1))
        return (-1e-3 * r2).cpu().tolist()

    smp = ConditionalSampler(dyn, timesteps=500)
    captures, init = [], _GraphedReverseStep.__init__

    def counting_init(self, *a, **k):
        captures.append(1)
        init(self, *a, **k)

    outs = []
    _GraphedReverseStep.__init__ = counting_init
    try:
        for graph in (False, True):
            torch.manual_seed(321)
            torch.cuda.manual_seed(321)
            xh, xp, lm, pm = smp.sample_given_pocket(pocket, sizes, timesteps=8, spsa=1, spsa_schedule=(5, 2), spsa_k=3,
                                                     reward_fn=reward, mixed_at=None, use_cuda_graph=graph)
            outs.append((xh.clone(), xp.clone()))
    finally:
        _GraphedReverseStep.__init__ = init
    assert len(captures) == 1                       # one capture, reused across the three SPSA updates
    scale = max(1.0, float(outs[0][0][:, :3].abs().max()))
    assert float((outs[0][0][:, :3] - outs[1][0][:, :3]).abs().max()) / scale < 5e-4
    assert float((outs[0][1] - outs[1][1]).abs().max()) / scale < 5e-4


# ---------------------------------------------------------------------------------------------------------------
# generate_ligands surface: PDB file -> pocket cache -> sampler -> GPU bond perception -> SDF
# ---------------------------------------------------------------------------------------------------------------
def _write_pdb(path, px, pt, lig_xyz):
    """Synthetic pocket as a PDB file: one GLY-named residue per 4 atoms plus a hetero ligand residue 900."""
    names = ['C', 'N', 'O', 'S']
    lines = []
    for i, (p, t) in enumerate(zip(px, pt)):
        el = names[int(t)] if int(t) < 4 else 'C'
        lines.append(f"ATOM  {i + 1:5d}  {el + str(i % 4):<3s} GLY A{i // 4 + 1:4d}    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}"
                     f"  1.00  0.00          {el:>2s}")
    for k, p in enumerate(lig_xyz):
        lines.append(f"HETATM{len(px) + k + 1:5d}  C{k:<2d} LIG A 900    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}  1.00  0.00           C")
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\nEND\n')


def test_generate_ligands_pdb_to_sdf(dyn, dev, tmp_path):
    from diffndm_b200 import ingest, output, synthetic
    from diffndm_b200.datasets import crossdock_dataset_info
    from diffndm_b200.generate import LigandGenerator
    from diffndm_b200.sampler import ConditionalSampler
    info = crossdock_dataset_info()
    px, pt = synthetic.synthetic_pocket(4, 120)
    px = (np.round(px, 3) + np.array([12.0, -30.0, 7.5], np.float32)).astype(np.float32)     # a pocket away from the origin
    pt = np.minimum(pt, 3)
    pdb = tmp_path / 'pocket.pdb'
    _write_pdb(pdb, px, pt, px.mean(0, keepdims=True) + np.array([[0, 0, 0], [1.4, 0, 0]], np.float32))
    hist = np.zeros((40, 200))
    hist[8:20, 100:130] = 1.0
    gen = LigandGenerator(ConditionalSampler(dyn, timesteps=500), info, size_histogram=hist)
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    mols, (xh_lig, xh_pocket, lig_mask, pocket_mask) = gen.generate_ligands(
        str(pdb), 4, ref_ligand='A:900', timesteps=20, n_nodes_min=9, return_tensors=True)
    assert len(mols) == 4 and gen.pockets.misses == 1
    sizes = torch.bincount(lig_mask).tolist()
    assert all(9 <= s < 20 for s in sizes) and [m.GetNumAtoms() for m in mols] == sizes
    # the pocket went back to where the PDB file has it (lightning_modules.py:918-925), the ligands sit inside it
    n_p = int(xh_pocket.shape[0]) // 4
    sel = ingest.pocket_arrays(ingest.get_pocket_from_ligand(ingest.parse_pdb(pdb), 'A:900'), info['atom_encoder'])[0]
    assert n_p == len(sel)
    assert np.abs(xh_pocket[:n_p, :3].cpu().numpy() - sel).max() < 2e-3
    com = sel.mean(0)
    assert all(np.linalg.norm(m.positions.mean(0) - com) < 15.0 for m in mols)
    # bonds of every molecule = the oracle's make_mol_edm on the same coordinates (bit-exact rule evaluation)
    x = xh_lig[:, :3].cpu().numpy()
    types = xh_lig[:, 3:].argmax(1).cpu().numpy()
    E_ref, _ = O.bond_orders(x, types, lig_mask.cpu().numpy(), np.asarray(info['bonds1'], np.float32),
                             np.asarray(info['bonds2'], np.float32), np.asarray(info['bonds3'], np.float32))[:2]
    for m, E in zip(mols, E_ref):
        ii, jj = np.nonzero(np.tril(E, -1))
        assert m.bonds.tolist() == np.stack([ii, jj, E[ii, jj]], 1).tolist()
    # second batch of the same pocket: served from the cache; script body writes the SDF
    n = gen.generate_to_sdf(str(pdb), tmp_path / 'out.sdf', n_samples=4, batch_size=2, num_nodes_lig=11, ref_ligand='A:900',
                            timesteps=10)
    assert n == 4 and gen.pockets.misses == 1 and gen.pockets.hits >= 2
    back = output.read_sdf(tmp_path / 'out.sdf')
    assert len(back) == 4 and all(1 <= m.GetNumAtoms() <= 11 for m in back)          # largest fragment of 11 atoms
    with pytest.raises(NotImplementedError):
        gen.generate_ligands(str(pdb), 2, ref_ligand='A:900', num_nodes_lig=torch.tensor([9, 9]), timesteps=2, sanitize=True)
    with pytest.raises(ValueError):
        LigandGenerator(gen.ddpm, info).generate_ligands(str(pdb), 2, ref_ligand='A:900', timesteps=2)


# ---------------------------------------------------------------------------------------------------------------
# (iv) full free-running trajectories: distribution-level comparison with the reference
# ---------------------------------------------------------------------------------------------------------------
def _ligand_stats(x_rel, types, sizes):
    """Per-ligand, translation-free statistics of end-of-trajectory ligands given relative to the pocket COM."""
    out = {'log_rg': [], 'log_com_dist': [], 'log_min_pair': []}
    off = 0
    for k in sizes:
        x = x_rel[off:off + k].astype(np.float64)
        off += k
        c = x.mean(0)
        out['log_rg'].append(np.log(np.sqrt(((x - c) ** 2).sum(1).mean())))
        out['log_com_dist'].append(np.log(np.linalg.norm(c) + 1e-9))
        d = np.sqrt(((x[:, None] - x[None]) ** 2).sum(-1))[np.triu_indices(k, 1)]
        out['log_min_pair'].append(np.log(d.min() + 1e-9))
    return {k: np.asarray(v) for k, v in out.items()}, np.bincount(types, minlength=10)


def test_trajectory_distribution_vs_reference(golden_weights, dev):
    """500-step free-running sampling, each side with its own Gaussian draws: the end-of-trajectory ligands of the CUDA
    engine and of the reference (``tests/golden/distribution.npz``, 64 ligands drawn by the unmodified reference on the CPU)
    must be statistically indistinguishable -- two-sample KS tests on per-ligand radius of gyration, distance to the pocket
    and closest atom pair, chi-square on the atom-type histogram; each at p > 1e-3 (bar stated here; with four tests the
    chance of a false alarm under equal distributions is < 0.4 %).  QED / SA need RDKit, which neither side has."""
    from scipy import stats
    from diffndm_b200.engine import B200EGNNDynamics
    from diffndm_b200.sampler import ConditionalSampler
    from diffndm_b200.weights import DynamicsConfig
    z = np.load(os.path.join(GOLDEN, 'distribution.npz'))
    assert int(z['weight_seed']) == 0 and abs(float(z['coord_head_gain']) - 0.3) < 1e-9      # = the golden_weights fixture
    sizes = z['sizes'].tolist()
    ref_stats, ref_hist = [], np.zeros(10, np.int64)
    for r in range(int(z['n_batches'])):
        s, h = _ligand_stats(z['x_rel'][r], z['types'][r], sizes)
        ref_stats.append(s)
        ref_hist += h
    ref = {k: np.concatenate([s[k] for s in ref_stats]) for k in ref_stats[0]}
    px, pt = z['pocket_x'], z['pocket_t']
    reps = 16                                                   # 16 x 16 = 256 ligands in one batch (samples are independent)
    all_sizes = np.tile(z['sizes'], reps)
    B, n_p = len(all_sizes), len(px)
    onehot = np.eye(10, dtype=np.float32)[pt]
    pocket = {'x': torch.from_numpy(np.tile(px, (B, 1))), 'one_hot': torch.from_numpy(np.tile(onehot, (B, 1))),
              'size': torch.tensor([n_p] * B), 'mask': torch.arange(B).repeat_interleave(n_p)}
    torch.manual_seed(77)
    torch.cuda.manual_seed(77)
    big = B200EGNNDynamics(DynamicsConfig(), golden_weights, max_nodes=B * (n_p + 16), max_edges=B * (n_p + 16) * 40,
                           max_samples=B).eval()
    smp = ConditionalSampler(big, timesteps=500)
    xh_l, xh_p, lm, pm = smp.sample_given_pocket(pocket, all_sizes, timesteps=500)
    pcom = torch.zeros((B, 3), device=xh_p.device).index_add_(0, pm, xh_p[:, :3]) / n_p
    x_rel = (xh_l[:, :3] - pcom[lm]).cpu().numpy()
    ours, our_hist = _ligand_stats(x_rel, xh_l[:, 3:].argmax(1).cpu().numpy(), all_sizes.tolist())
    report = {}
    for k in ref:
        assert np.isfinite(ours[k]).all()
        report[k] = stats.ks_2samp(ref[k], ours[k]).pvalue
    keep = (ref_hist + our_hist) > 0
    report['atom_types'] = stats.chi2_contingency(np.stack([ref_hist[keep], our_hist[keep]]))[1]
    print('distribution p-values:', {k: round(float(v), 4) for k, v in report.items()},
          'medians ref/ours:', {k: (round(float(np.median(ref[k])), 3), round(float(np.median(ours[k])), 3)) for k in ref})
    assert all(p > 1e-3 for p in report.values()), report


def test_inpaint_ligand_keeps_the_fixed_substructure(dyn, dev, tmp_path):
    """``LigandGenerator.inpaint_ligand`` (inpaint.py:65-188 surface): the atoms named in ``fix_atoms`` keep their element
    and their geometry (up to the p(x|z_0) head's noise), every sample gets n_fixed + add_n_nodes atoms, and the output is
    moved back into the frame of the PDB file."""
    from diffndm_b200 import synthetic
    from diffndm_b200.datasets import crossdock_dataset_info
    from diffndm_b200.generate import LigandGenerator
    from diffndm_b200.sampler import ConditionalSampler
    info = crossdock_dataset_info()
    px, pt = synthetic.synthetic_pocket(6, 100)
    px = (np.round(px, 3) + np.array([-8.0, 21.0, 3.5], np.float32)).astype(np.float32)
    frag = px.mean(0, keepdims=True) + np.array([[0, 0, 0], [1.4, 0, 0], [2.1, 1.2, 0]], np.float32)
    pdb = tmp_path / 'pocket.pdb'
    _write_pdb(pdb, px, np.minimum(pt, 3), frag)
    gen = LigandGenerator(ConditionalSampler(dyn, timesteps=500), info)
    torch.manual_seed(11)
    torch.cuda.manual_seed(11)
    mols, (xh_lig, xh_pocket, lig_mask, pocket_mask) = gen.inpaint_ligand(
        str(pdb), 3, 'A:900', ['C0', 'C2'], add_n_nodes=5, timesteps=25, resamplings=2, return_tensors=True)
    assert len(mols) == 3 and torch.bincount(lig_mask).tolist() == [7, 7, 7]
    d_ref = float(np.linalg.norm(frag[0] - frag[2]))
    for b in range(3):
        x = xh_lig[lig_mask == b][:, :3].cpu().numpy()
        t = xh_lig[lig_mask == b][:, 3:].argmax(1).cpu().numpy()
        assert t[0] == 0 and t[1] == 0                                          # both fixed atoms are carbons
        assert abs(np.linalg.norm(x[0] - x[1]) - d_ref) < 0.3                   # fragment geometry kept
        assert np.abs(x[:2] - frag[[0, 2]]).max() < 0.5                         # and it sits where the PDB file has it
    with pytest.raises(ValueError):
        gen.inpaint_ligand(str(pdb), 2, 'A:900', ['C0'], timesteps=5)            # no size prior, no add_n_nodes


def test_prefiltered_reward_scores_only_plausible_candidates(dyn, dev):
    """``chem.PrefilteredReward``: candidates without a bonded majority fragment never reach the host scorer and get the
    rejected score; the others are scored in their original order."""
    from diffndm_b200.chem import BondPerception, PrefilteredReward
    from diffndm_b200.datasets import crossdock_dataset_info
    info = crossdock_dataset_info()
    chain = np.array([[0, 0, 0], [1.5, 0, 0], [2.2, 1.3, 0], [3.7, 1.3, 0]], np.float32)         # C-C-C-C, all bonded
    gas = np.array([[0, 0, 0], [6, 0, 0], [0, 6, 0], [0, 0, 6]], np.float32)                       # four lone atoms
    x = torch.from_numpy(np.concatenate([gas, chain, gas + 1.0, chain + 2.0])).to(dev)
    types = torch.zeros(16, dtype=torch.long, device=dev)
    mask = torch.arange(4, device=dev).repeat_interleave(4)
    seen = []

    def scorer(xs, ts, ms):
        seen.append((xs.cpu().numpy().copy(), ms.cpu().numpy().copy()))
        return [10.0 + float(xs[ms == i][:, 0].mean()) for i in range(int(ms.max()) + 1)]

    pre = PrefilteredReward(scorer, BondPerception(dyn.engine, info), rejected_score=-1.0)
    out = pre(x, types, mask)
    assert out[0] == -1.0 and out[2] == -1.0 and pre.rejected == 2 and pre.scored == 2
    assert abs(out[1] - (10.0 + chain[:, 0].mean())) < 1e-5 and abs(out[3] - (12.0 + chain[:, 0].mean())) < 1e-5
    assert len(seen) == 1 and seen[0][1].tolist() == [0] * 4 + [1] * 4                               # compacted ids
    assert np.allclose(seen[0][0], np.concatenate([chain, chain + 2.0]))
    again = pre(x[4:8], types[4:8], mask[:4])                                                        # nothing rejected: pass-through
    assert len(again) == 1 and abs(again[0] - (10.0 + float(chain[:, 0].mean()))) < 1e-5 and len(seen) == 2


def test_spsa_overlapped_scoring_matches_plain(dyn, dev):
    """With a scorer that offers ``submit`` (hostpool.PooledReward) an SPSA round is split into its +U and -U halves: the
    first half is scored by the worker processes while the GPU denoises the second.  Same perturbations and noise in,
    same update out (the halves see the denoiser at another batch composition: tolerance, not bit equality)."""
    from diffndm_b200 import synthetic
    from diffndm_b200.hostpool import PooledReward, radius_of_gyration_score
    from diffndm_b200.sampler import ConditionalSampler
    px, pt = synthetic.synthetic_pocket(8, 70)
    sizes = np.array([6, 11, 8, 9])
    b = synthetic.make_batch(px, pt, sizes, 8)
    k, B = 4, 4
    rng = np.random.default_rng(3)
    U = torch.from_numpy((1e-3 * rng.standard_normal((k, len(b['lig_mask']), 3))).astype(np.float32))
    x0_noise = torch.from_numpy(rng.standard_normal((2 * k, len(b['lig_mask']), 13)).astype(np.float32)).to(dev)
    smp = ConditionalSampler(dyn, timesteps=500)
    t_arr = torch.full((B, 1), 16 / 500)
    args = (_t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), _t(b['lig_mask'], dev), _t(b['pocket_mask'], dev), t_arr, B, 1e-3)
    with PooledReward(radius_of_gyration_score, workers=2, chunk=3) as pool:
        submits = []
        orig = pool.submit
        pool.submit = lambda *a, **kw: (submits.append(kw.get('after') is not None), orig(*a, **kw))[1]
        z1, p1 = smp.my_update_z_lig(*args, pool, guidance_scale=1e-3, k=k, perturbations=U, x0_noise=x0_noise)
        assert submits == [True, True]                                       # two halves, each gated by its CUDA event
        smp.overlap_scoring = False
        z2, p2 = smp.my_update_z_lig(*args, pool, guidance_scale=1e-3, k=k, perturbations=U, x0_noise=x0_noise)
        assert len(submits) == 3 and submits[2] is False                     # plain path: one blocking call
    scale = max(1.0, float(z2.abs().max()))
    assert float((z1 - z2).abs().max()) < 2e-4 * scale and float((p1 - p2).abs().max()) < 2e-4 * scale


def test_atp_overlapped_scoring_matches_plain(dyn, dev):
    """ATP event with a scorer that offers ``submit``: the current candidates are scored by the worker processes while
    the GPU runs their x0 look-ahead.  Same random draws in, same winners out."""
    from diffndm_b200 import synthetic
    from diffndm_b200.hostpool import PooledReward, radius_of_gyration_score
    from diffndm_b200.sampler import ConditionalSampler
    px, pt = synthetic.synthetic_pocket(8, 50)
    sizes = np.array([6, 9, 5, 7])
    b = synthetic.make_batch(px, pt, sizes, 8)
    B, G, s = 4, 3, 20
    smp = ConditionalSampler(dyn, timesteps=500)
    s_arr, t_arr = torch.full((B, 1), s / 500), torch.full((B, 1), (s + 1) / 500)
    args = (s, s_arr, t_arr, _t(b['xh_lig'], dev), _t(b['xh_pocket'], dev), _t(b['lig_mask'], dev), _t(b['pocket_mask'], dev), B)
    outs = []
    with PooledReward(radius_of_gyration_score, workers=2) as pool:
        gated = []
        orig = pool.submit
        pool.submit = lambda *a, **kw: (gated.append(kw.get('after') is not None), orig(*a, **kw))[1]
        for overlap in (True, False):
            smp.overlap_scoring = overlap
            torch.manual_seed(3)
            torch.cuda.manual_seed(3)
            outs.append(smp._atp_event(*args, pool, G))
    assert gated.count(True) == 1                                            # one event-gated submit in the overlapped run
    (z1, p1, m1), (z2, p2, m2) = outs
    assert torch.equal(m1, m2) and torch.equal(z1, z2) and torch.equal(p1, p2)


def test_generate_ligands_script(dev, tmp_path):
    """scripts/generate_ligands.py: the reference script's flags on the B200 engine (random weights: no checkpoint here)."""
    import importlib.util
    from diffndm_b200 import output, synthetic
    spec = importlib.util.spec_from_file_location('b200_generate_ligands',
                                                  os.path.join(os.path.dirname(GOLDEN), '..', 'scripts', 'generate_ligands.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    px, pt = synthetic.synthetic_pocket(12, 90)
    px = np.round(px, 3).astype(np.float32)
    pdb, sdf = tmp_path / 'p.pdb', tmp_path / 'o.sdf'
    _write_pdb(pdb, px, np.minimum(pt, 3), px.mean(0, keepdims=True) + np.array([[0, 0, 0], [1.4, 0, 0]], np.float32))
    n = mod.main(['--random_init', '0', '--pdbfile', str(pdb), '--ref_ligand', 'A:900', '--outfile', str(sdf), '--n_samples', '4',
                  '--batch_size', '2', '--num_nodes_lig', '10', '--timesteps', '8', '--seed', '1', '--all_frags'])
    mols = output.read_sdf(sdf)
    assert n == 4 and len(mols) == 4 and all(m.GetNumAtoms() == 10 for m in mols)
    with pytest.raises(SystemExit):
        mod.main(['--random_init', '0', '--pdbfile', str(pdb), '--ref_ligand', 'A:900', '--outfile', str(sdf), '--SPSA', '1'])


def test_inpaint_script(dev, tmp_path):
    """scripts/inpaint.py: the reference script's flags; the fixed fragment comes from an SDF file here."""
    import importlib.util
    from diffndm_b200 import output, synthetic
    spec = importlib.util.spec_from_file_location('b200_inpaint', os.path.join(os.path.dirname(GOLDEN), '..', 'scripts', 'inpaint.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    px, pt = synthetic.synthetic_pocket(13, 90)
    px = np.round(px, 3).astype(np.float32)
    frag = px.mean(0, keepdims=True) + np.array([[0, 0, 0], [1.4, 0, 0], [2.1, 1.2, 0]], np.float32)
    pdb, frag_sdf, sdf = tmp_path / 'p.pdb', tmp_path / 'frag.sdf', tmp_path / 'o.sdf'
    _write_pdb(pdb, px, np.minimum(pt, 3), frag)
    output.write_sdf_file(frag_sdf, [output.Molecule(['C', 'N'], frag[:2], np.array([[1, 0, 1]]))])
    n = mod.main(['--random_init', '0', '--pdbfile', str(pdb), '--ref_ligand', 'A:900', '--fix_atoms', str(frag_sdf), '--outfile',
                  str(sdf), '--n_samples', '3', '--add_n_nodes', '4', '--timesteps', '12', '--resamplings', '2', '--seed', '2'])
    mols = output.read_sdf(sdf)
    assert n == 3 and len(mols) == 3
    for m in mols:
        assert m.GetNumAtoms() == 6 and m.symbols[:2] == ['C', 'N']
        assert np.abs(m.positions[:2] - frag[:2]).max() < 0.5
