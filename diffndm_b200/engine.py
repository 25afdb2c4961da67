"""ctypes binding of ``libdiffndm_b200.so`` (include/diffndm_b200.h) and the host-side mirror of the
reference operator surface.

``B200EGNNDynamics`` is a drop-in for ``EGNNDynamics`` (reference dynamics.py:10-167) in the conditional
sampler: same call signature ``forward(xh_atoms, xh_residues, t, mask_atoms, mask_residues)``, same outputs,
same error behaviour (``ValueError("NaN detected in EGNN output")`` in eval mode, dynamics.py:155-159).  It is
installed with ``model.ddpm.dynamics = B200EGNNDynamics.from_reference(model.ddpm.dynamics)``.

There is no CPU or PyTorch fallback: a missing library or a non-CUDA input raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from .weights import COMPILED_HIDDEN_NF, DynamicsConfig, engine_table, expected_keys

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libdiffndm_b200.so')

FLAG_NAN = 1
FLAG_COM_DRIFT = 2
FLAG_EDGE_OVERFLOW = 4
FLAG_MOL_TOO_LARGE = 8

EXPORTED_SYMBOLS = [
    'dndm_version', 'dndm_last_error', 'dndm_launch_count', 'dndm_engine_create', 'dndm_engine_destroy', 'dndm_engine_load_weights',
    'dndm_egnn_forward', 'dndm_radius_graph', 'dndm_sampler_step', 'dndm_read_flags', 'dndm_debug_copy',
    'dndm_set_trace', 'dndm_set_profile', 'dndm_get_profile', 'dndm_set_static_masks', 'dndm_bond_orders', 'dndm_test_gemm',
]


class DndmConfig(ctypes.Structure):
    _fields_ = [
        ('atom_nf', ctypes.c_int32), ('residue_nf', ctypes.c_int32), ('joint_nf', ctypes.c_int32),
        ('hidden_nf', ctypes.c_int32), ('n_layers', ctypes.c_int32),
        ('edge_cutoff_ligand', ctypes.c_float), ('edge_cutoff_pocket', ctypes.c_float),
        ('edge_cutoff_interaction', ctypes.c_float), ('norm_constant', ctypes.c_float),
        ('normalization_factor', ctypes.c_float), ('coords_range', ctypes.c_float),
        ('max_nodes', ctypes.c_int32), ('max_edges', ctypes.c_int32), ('max_samples', ctypes.c_int32),
        ('device', ctypes.c_int32),
    ]


class DndmWeight(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char_p), ('data', ctypes.POINTER(ctypes.c_float)),
                ('rows', ctypes.c_int32), ('cols', ctypes.c_int32)]


_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen the C-ABI library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(f'{_LIB_PATH} is missing: run `python -m diffndm_b200.build` (or __graft_entry__.build()). '
                           'diffndm_b200 has no CPU / PyTorch fallback for the denoiser.')
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    lib.dndm_version.restype = ctypes.c_char_p
    lib.dndm_last_error.restype = ctypes.c_char_p
    lib.dndm_launch_count.restype = ctypes.c_int64
    lib.dndm_engine_create.argtypes = [ctypes.POINTER(DndmConfig), ctypes.POINTER(vp)]
    lib.dndm_engine_destroy.argtypes = [vp]
    lib.dndm_engine_destroy.restype = None
    lib.dndm_engine_load_weights.argtypes = [vp, ctypes.POINTER(DndmWeight), i32]
    lib.dndm_egnn_forward.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.dndm_radius_graph.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, i32, ctypes.POINTER(i32), vp]
    lib.dndm_sampler_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, f32, vp, vp, i32, i32, i32, vp, vp, i32, vp]
    lib.dndm_read_flags.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32), vp]
    lib.dndm_debug_copy.argtypes = [vp, i32, vp, i64, vp]
    lib.dndm_debug_copy.restype = i64
    lib.dndm_set_trace.argtypes = [vp, vp, vp, i32]
    lib.dndm_set_profile.argtypes = [vp, i32]
    lib.dndm_set_static_masks.argtypes = [vp, i32]
    lib.dndm_bond_orders.argtypes = [vp, vp, i32, vp, vp, i32, i32, vp, vp, vp, i32, f32, f32, f32, vp, vp, i64, vp, vp, vp]
    lib.dndm_get_profile.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32), i32]
    lib.dndm_test_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    _lib = lib
    return lib


def _check(lib, rc, what):
    if rc < 0:
        raise RuntimeError(f'{what} failed ({rc}): {lib.dndm_last_error().decode()}')
    return rc


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: the B200 engine has no CPU path')
    return t.contiguous().float()


class Engine:
    """Thin owner of one ``DndmEngine*``.  Not re-entrant; all work is enqueued on torch's current stream."""

    def __init__(self, cfg: DynamicsConfig = DynamicsConfig(), max_nodes: int = 40960, max_edges: int = 1 << 20,
                 max_samples: int = 256, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError('diffndm_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = load_library()
        self.cfg = cfg
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.max_nodes, self.max_edges, self.max_samples = int(max_nodes), int(max_edges), int(max_samples)
        if cfg.update_pocket_coords or not cfg.condition_time or cfg.reflection_equivariant \
                or not cfg.attention or not cfg.tanh or cfg.inv_sublayers != 1:
            raise NotImplementedError('engine implements the pocket-conditional denoisers: attention, tanh, cross-product '
                                      'MLP, time conditioning, frozen pocket')
        if cfg.hidden_nf > COMPILED_HIDDEN_NF:
            raise NotImplementedError(f'hidden_nf = {cfg.hidden_nf}: the kernels are compiled for {COMPILED_HIDDEN_NF} hidden '
                                      'channels; narrower networks are zero-padded (weights.engine_table), wider ones are not built')
        # the device engine always runs at the compiled width; edge-type embeddings arrive folded (weights.engine_table)
        c = DndmConfig(cfg.atom_nf, cfg.residue_nf, cfg.joint_nf, COMPILED_HIDDEN_NF, cfg.n_layers,
                       -1.0 if cfg.edge_cutoff_ligand is None else cfg.edge_cutoff_ligand,
                       -1.0 if cfg.edge_cutoff_pocket is None else cfg.edge_cutoff_pocket,
                       -1.0 if cfg.edge_cutoff_interaction is None else cfg.edge_cutoff_interaction,
                       cfg.norm_constant, cfg.normalization_factor, cfg.coords_range,
                       self.max_nodes, self.max_edges, self.max_samples, self.device)
        h = ctypes.c_void_p()
        _check(self.lib, self.lib.dndm_engine_create(ctypes.byref(c), ctypes.byref(h)), 'dndm_engine_create')
        self._h = h
        self._trace = None

    def close(self):
        if getattr(self, '_h', None):
            self.lib.dndm_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights -----------------------------------------------------------------------------------
    weights_version = 0          # bumped by load_weights: captured CUDA graphs hold pointers into the packed weights

    def load_weights(self, state: Mapping[str, object]):
        """``state``: reference ``EGNNDynamics.state_dict()`` (tensors) or a dict of numpy arrays, in the shapes of
        ``self.cfg`` (any hidden_nf <= 256, with or without the edge-type embedding); the device engine receives the
        table ``weights.engine_table`` derives from it (zero-padded to the compiled width, embedding folded)."""
        self.weights_version += 1
        host = {}
        for name, shape in expected_keys(self.cfg):
            if name not in state:
                raise KeyError(f'missing weight {name}')
            v = state[name]
            a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            a = np.ascontiguousarray(a, dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f'{name}: shape {a.shape} != expected {shape}')
            host[name] = a
        _, table = engine_table(self.cfg, host)
        keep, arr = [], (DndmWeight * len(table))()
        for i, (name, a) in enumerate(table.items()):
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            rows, cols = (a.shape[0], a.shape[1]) if a.ndim == 2 else (a.shape[0], 1)
            arr[i] = DndmWeight(name.encode(), a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), rows, cols)
        _check(self.lib, self.lib.dndm_engine_load_weights(self._h, arr, len(keep)), 'dndm_engine_load_weights')

    # -- hot path ------------------------------------------------------------------------------------
    def forward(self, xh_lig, xh_pocket, t, lig_mask, pocket_mask, n_samples: int, out_lig=None, out_pocket=None,
                want_pocket: bool = True):
        xh_lig = _dev_f32(xh_lig, 'xh_atoms')
        xh_pocket = _dev_f32(xh_pocket, 'xh_residues')
        t = _dev_f32(t, 't').reshape(-1)
        lig_mask = lig_mask.contiguous().long()
        pocket_mask = pocket_mask.contiguous().long()
        if out_lig is None:
            out_lig = torch.empty_like(xh_lig)
        if out_pocket is None and want_pocket:
            out_pocket = torch.empty_like(xh_pocket)
        rc = self.lib.dndm_egnn_forward(self._h, _ptr(xh_lig), _ptr(xh_pocket), _ptr(t), t.numel(), _ptr(lig_mask),
                                        _ptr(pocket_mask), xh_lig.shape[0], xh_pocket.shape[0], int(n_samples),
                                        _ptr(out_lig), _ptr(out_pocket), _stream())
        _check(self.lib, rc, 'dndm_egnn_forward')
        return out_lig, out_pocket

    def radius_graph(self, xh_lig, xh_pocket, lig_mask, pocket_mask, n_samples: int):
        """Returns (row_ptr int32 [N+1], col int32 [E]) device tensors -- dynamics.py:169-187 as CSR."""
        xh_lig = _dev_f32(xh_lig, 'xh_atoms')
        xh_pocket = _dev_f32(xh_pocket, 'xh_residues')
        lig_mask = lig_mask.contiguous().long()
        pocket_mask = pocket_mask.contiguous().long()
        n = xh_lig.shape[0] + xh_pocket.shape[0]
        row_ptr = torch.empty(n + 1, dtype=torch.int32, device=xh_lig.device)
        col = torch.empty(self.max_edges, dtype=torch.int32, device=xh_lig.device)
        ne = ctypes.c_int32(0)
        rc = self.lib.dndm_radius_graph(self._h, _ptr(xh_lig), _ptr(xh_pocket), _ptr(lig_mask), _ptr(pocket_mask),
                                        xh_lig.shape[0], xh_pocket.shape[0], int(n_samples), _ptr(row_ptr), _ptr(col),
                                        self.max_edges, ctypes.byref(ne), _stream())
        _check(self.lib, rc, 'dndm_radius_graph')
        return row_ptr, col[:ne.value]

    def sampler_step(self, z_in, eps, noise, xh_pocket, coef, lig_mask, pocket_mask, n_samples: int, grad=None,
                     lam: float = 0.0, z_out=None, pocket_out=None, check_com: bool = False):
        z_in = _dev_f32(z_in, 'z')
        noise = _dev_f32(noise, 'noise')
        xh_pocket = _dev_f32(xh_pocket, 'xh_pocket')
        coef = _dev_f32(coef, 'coef')
        eps = None if eps is None else _dev_f32(eps, 'eps')
        grad = None if grad is None else _dev_f32(grad, 'grad')
        lig_mask = lig_mask.contiguous().long()
        pocket_mask = pocket_mask.contiguous().long()
        if z_in.shape[1] != 3 + self.cfg.atom_nf or xh_pocket.shape[1] != 3 + self.cfg.residue_nf:
            raise ValueError(f'sampler_step: rows must be [3 + {self.cfg.atom_nf}] (ligand) and [3 + {self.cfg.residue_nf}] (pocket) wide, '
                             f'got {z_in.shape[1]} and {xh_pocket.shape[1]}')
        if z_out is None:
            z_out = torch.empty_like(z_in)
        if pocket_out is None:
            pocket_out = torch.empty_like(xh_pocket)
        rc = self.lib.dndm_sampler_step(self._h, _ptr(z_in), _ptr(eps), _ptr(noise), _ptr(xh_pocket), _ptr(coef), _ptr(grad),
                                        ctypes.c_float(lam), _ptr(lig_mask), _ptr(pocket_mask), z_in.shape[0],
                                        xh_pocket.shape[0], int(n_samples), _ptr(z_out), _ptr(pocket_out), int(bool(check_com)),
                                        _stream())
        _check(self.lib, rc, 'dndm_sampler_step')
        return z_out, pocket_out

    _pending_flags = 0

    def read_flags(self, consume: int = 0xFFFFFFFF) -> int:
        """Sticky device flag word (NaN / COM drift / edge overflow / molecule too large), read and cleared on the device.
        ``consume``: the bits the caller handles; the others stay pending on the host and are returned again by the next
        call, so that a caller that only looks at NaN does not swallow a COM-drift bit meant for the sampler."""
        f = ctypes.c_uint32(0)
        _check(self.lib, self.lib.dndm_read_flags(self._h, ctypes.byref(f), _stream()), 'dndm_read_flags')
        flags = int(f.value) | self._pending_flags
        self._pending_flags = flags & ~consume & 0xFFFFFFFF
        return flags

    # -- introspection (tests / bench) ---------------------------------------------------------------
    def set_trace(self, max_nodes: int):
        dev = torch.device('cuda', self.device)
        L = self.cfg.n_layers
        self._trace = (torch.zeros(L, max_nodes, 256, device=dev), torch.zeros(L, max_nodes, 3, device=dev))
        _check(self.lib, self.lib.dndm_set_trace(self._h, _ptr(self._trace[0]), _ptr(self._trace[1]), max_nodes), 'set_trace')
        return self._trace

    def clear_trace(self):
        self._trace = None
        self.lib.dndm_set_trace(self._h, None, None, 0)

    PROFILE_CATEGORIES = ('gcl_edge_kernel', 'head_edge_kernel', 'node_gemm', 'radius_graph', 'node_other')

    def set_static_masks(self, on: bool):
        """Promise that mask tensors passed again (same storage, same sizes) still hold the same values; the engine then
        skips re-deriving the per-sample offsets.  Turn it off (or toggle it) whenever a mask tensor is rewritten."""
        _check(self.lib, self.lib.dndm_set_static_masks(self._h, int(bool(on))), 'dndm_set_static_masks')

    def set_profile(self, on: bool):
        _check(self.lib, self.lib.dndm_set_profile(self._h, int(on)), 'dndm_set_profile')

    def get_profile(self):
        """{category: (total_ms, timed_sections)} since the last call (synchronises the device)."""
        n = len(self.PROFILE_CATEGORIES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int32 * n)()
        _check(self.lib, self.lib.dndm_get_profile(self._h, ms, cnt, n), 'dndm_get_profile')
        return {c: (ms[i], cnt[i]) for i, c in enumerate(self.PROFILE_CATEGORIES)}

    def graph_stats(self):
        """(E, E_ligand_receiver) of the last forward / radius_graph call (synchronises)."""
        return self.graph_stats_full()[:2]

    def debug_h0(self, n_nodes: int):
        """Encoder + embedding output h_0 [n_nodes, 256] of the last forward (kept only while a trace is set)."""
        t = torch.empty((n_nodes, 256), dtype=torch.float32, device=torch.device('cuda', self.device))
        _check(self.lib, self.lib.dndm_debug_copy(self._h, 8, _ptr(t), t.numel() * 4, _stream()), 'dndm_debug_copy')
        return t

    def pocket_list_state(self):
        """Introspection for tests: int32 [8] state of the pocket-pocket candidate lists (include/diffndm_b200.h, buffer 7)."""
        t = torch.zeros(8, dtype=torch.int32, device=torch.device('cuda', self.device))
        _check(self.lib, self.lib.dndm_debug_copy(self._h, 7, _ptr(t), 32, _stream()), 'dndm_debug_copy')
        return t.cpu().tolist()

    def graph_stats_full(self):
        """(E, E_ligand_receiver, E_last_block): the third entry is the number of edges the last block aggregates when
        the pocket output is not requested (ligand receivers + their pocket senders); stale otherwise."""
        t = torch.zeros(4, dtype=torch.int32, device=torch.device('cuda', self.device))
        _check(self.lib, self.lib.dndm_debug_copy(self._h, 4, _ptr(t), 16, _stream()), 'dndm_debug_copy')
        e, el, ea, _ = t.cpu().tolist()
        return e, el, ea

    def graph_stats_pruned(self):
        """(E, E_ligand_receiver, [E of the last block, of the block before it, ...]) of the last forward without pocket output:
        the edges the trailing blocks aggregate over (exact dead-work elimination, csrc/graph.cuh; DNDM_PRUNE_LEVELS of them
        -- default 2, at most 3 and at most n_layers); stale otherwise."""
        t = torch.zeros(8, dtype=torch.int32, device=torch.device('cuda', self.device))
        _check(self.lib, self.lib.dndm_debug_copy(self._h, 4, _ptr(t), 32, _stream()), 'dndm_debug_copy')
        v = t.cpu().tolist()
        levels = max(1, min(int(os.environ.get('DNDM_PRUNE_LEVELS', 2)), 3, self.cfg.n_layers))
        return v[0], v[1], v[2:2 + levels]


def launch_count() -> int:
    """Kernels of libdiffndm_b200 launched (or captured) by this process so far."""
    return int(load_library().dndm_launch_count())


def test_gemm(a_bf16: torch.Tensor, w_bf16: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = 0, bn: int = 256,
              residual: Optional[torch.Tensor] = None, n_tail_groups: int = 0, m_tail: int = 0, want_bf16: bool = False):
    """C = A W^T (+bias)(+residual)(SiLU) through ``gemm_pair_kernel``, the weight-resident tcgen05 node GEMM, with the
    forward's launch geometry.  Returns (out_f32 [M,N], out_bf16 [M,N] or None); outputs the tail groups do not cover are 0."""
    lib = load_library()
    M, K = a_bf16.shape
    N = w_bf16.shape[0]
    rows = -(-M // 128) * 128
    a_pad = torch.zeros(rows, K, dtype=torch.bfloat16, device=a_bf16.device)
    a_pad[:M] = a_bf16
    out = torch.zeros(M, N, dtype=torch.float32, device=a_bf16.device)
    out16 = torch.zeros(rows, N, dtype=torch.bfloat16, device=a_bf16.device) if want_bf16 else None
    rc = lib.dndm_test_gemm(_ptr(a_pad), _ptr(w_bf16.contiguous()), _ptr(bias), _ptr(residual), act, M, N, K, bn, n_tail_groups,
                            m_tail, _ptr(out), _ptr(out16), _stream())
    _check(lib, rc, 'dndm_test_gemm')
    return out, (None if out16 is None else out16[:M])


class B200EGNNDynamics(torch.nn.Module):
    """Host-side mirror of ``EGNNDynamics`` (reference dynamics.py:10-167) backed by the CUDA engine."""

    def __init__(self, cfg: DynamicsConfig = DynamicsConfig(), state: Optional[Mapping[str, object]] = None,
                 max_nodes: int = 40960, max_edges: int = 1 << 20, max_samples: int = 256, check_nan: bool = True):
        super().__init__()
        self.cfg = cfg
        self.engine = Engine(cfg, max_nodes, max_edges, max_samples)
        self.update_pocket_coords = cfg.update_pocket_coords     # read by ConditionalDDPM.__init__ (conditional_model.py:24)
        self.n_dims = cfg.n_dims
        self.check_nan = check_nan          # the reference syncs on NaN every call (dynamics.py:155); can be deferred
        self.compute_pocket_output = True   # every conditional caller discards it (`eps, _ = ...`)
        if state is not None:
            self.engine.load_weights(state)

    @classmethod
    def from_reference(cls, ref_module, **kw) -> 'B200EGNNDynamics':
        """Build from a live reference ``EGNNDynamics`` (weights packed once; call ``reload`` after load_state_dict)."""
        egnn = ref_module.egnn
        blk = egnn._modules['e_block_0']
        cfg = DynamicsConfig(
            atom_nf=ref_module.atom_encoder[0].in_features, residue_nf=ref_module.residue_encoder[0].in_features,
            n_dims=ref_module.n_dims, joint_nf=ref_module.atom_encoder[2].out_features, hidden_nf=egnn.hidden_nf,
            n_layers=egnn.n_layers, edge_cutoff_ligand=ref_module.edge_cutoff_l, edge_cutoff_pocket=ref_module.edge_cutoff_p,
            edge_cutoff_interaction=ref_module.edge_cutoff_i, norm_constant=blk.norm_constant,
            normalization_factor=egnn.normalization_factor, coords_range=blk.coords_range_layer,
            attention=blk._modules['gcl_0'].attention, tanh=blk._modules['gcl_equiv'].tanh,
            reflection_equivariant=egnn.reflection_equiv, inv_sublayers=blk.n_layers,
            edge_embedding_dim=(ref_module.edge_nf or None), update_pocket_coords=ref_module.update_pocket_coords,
            condition_time=ref_module.condition_time)
        return cls(cfg, ref_module.state_dict(), **kw)

    def reload(self, state):
        self.engine.load_weights(state)

    def forward(self, xh_atoms, xh_residues, t, mask_atoms, mask_residues, n_samples: Optional[int] = None):
        if n_samples is None:
            # t is [B,1] in every sampler call site (conditional_model.py:946-949); a scalar t needs the mask
            n_samples = int(t.numel()) if t.numel() > 1 else int(max(int(mask_atoms.max()), int(mask_residues.max()))) + 1
        out_l, out_p = self.engine.forward(xh_atoms, xh_residues, t, mask_atoms, mask_residues, n_samples,
                                           want_pocket=self.compute_pocket_output)
        if self.check_nan:
            flags = self.engine.read_flags(consume=FLAG_NAN | FLAG_EDGE_OVERFLOW)
            if flags & FLAG_EDGE_OVERFLOW:
                raise RuntimeError('diffndm_b200: edge capacity exceeded (raise max_edges)')
            if flags & FLAG_NAN:
                if not self.training:
                    raise ValueError("NaN detected in EGNN output")
                out_l[:, :3] = torch.nan_to_num(out_l[:, :3], nan=0.0)           # vel[isnan] = 0, dynamics.py:156-157
        return out_l, out_p
