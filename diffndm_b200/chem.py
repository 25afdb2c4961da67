"""Bond perception on the GPU for pre-filtering candidate molecules.

Mirrors the numeric part of the reference's ``make_mol_edm`` / ``get_bond_order_batch``
(analysis/molecule_builder.py:30-55, 100-113): the directed lower-triangular bond-order matrix of every molecule of a
batch from the distance tables in ``dataset_info`` -- plus what a filter needs from it (valences, valence violations,
fragments), so that obviously broken candidates of an SPSA / ATP round can be dropped before the host pays for
OpenBabel + RDKit.  RDKit scoring itself stays on the host, exactly as in the reference.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Mapping, Optional, Sequence

import torch

from .engine import Engine, _check, _ptr, _stream

# constants.py:17 and :19-22 of the reference (margins in pm; maximum valence per element, lists -> their maximum)
MARGINS = (3.0, 2.0, 1.0)
ALLOWED_BONDS = {'H': 1, 'C': 4, 'N': 3, 'O': 2, 'F': 1, 'B': 3, 'Al': 3, 'Si': 4, 'P': 5, 'S': 4, 'Cl': 1, 'As': 3,
                 'Br': 1, 'I': 1, 'Hg': 2, 'Bi': 5}


class BondPerception:
    """Device copies of one ``dataset_info``'s tables; ``__call__`` runs ``dndm_bond_orders``."""

    def __init__(self, engine: Engine, dataset_info: Mapping[str, object], margins: Sequence[float] = MARGINS):
        self.engine = engine
        dev = torch.device('cuda', engine.device)
        self.b1, self.b2, self.b3 = (torch.tensor(dataset_info[k], dtype=torch.float32, device=dev).contiguous()
                                     for k in ('bonds1', 'bonds2', 'bonds3'))
        self.n_types = int(self.b1.shape[0])
        dec = dataset_info.get('atom_decoder')
        self.allowed = None
        if dec is not None:
            self.allowed = torch.tensor([ALLOWED_BONDS.get(a, 1 << 20) for a in dec][:self.n_types], dtype=torch.int32,
                                        device=dev)
        self.margins = tuple(float(m) for m in margins)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, atom_types: torch.Tensor, mol_mask: torch.Tensor, n_mols: int,
                 return_matrices: bool = True) -> Dict[str, object]:
        """x [N,>=3] fp32 (Angstrom), atom_types [N] int64 decoder indices, mol_mask [N] sorted int64.
        Returns valence [N] int32, n_bonds / n_components / largest_component / valence_violations [n_mols] int32 and
        (optionally) ``E``: a list of int8 [n_b, n_b] lower-triangular bond-order matrices like make_mol_edm's ``E``."""
        eng = self.engine
        dev = torch.device('cuda', eng.device)
        if x.device != dev or x.dtype != torch.float32:
            raise ValueError('x must be a float32 CUDA tensor on the engine device')
        x = x.contiguous()
        atom_types = atom_types.to(dev).long().contiguous()
        mol_mask = mol_mask.to(dev).long().contiguous()
        n = int(x.shape[0])
        sizes = torch.bincount(mol_mask, minlength=n_mols)
        valence = torch.empty(n, dtype=torch.int32, device=dev)
        stats = torch.empty((n_mols, 4), dtype=torch.int32, device=dev)
        e_flat, cap = None, 0
        if return_matrices:
            cap = int((sizes.long() ** 2).sum().item())
            e_flat = torch.empty(max(cap, 1), dtype=torch.int8, device=dev)
        m1, m2, m3 = self.margins
        rc = eng.lib.dndm_bond_orders(eng._h, _ptr(x), int(x.shape[1]), _ptr(atom_types), _ptr(mol_mask), n, int(n_mols),
                                      _ptr(self.b1), _ptr(self.b2), _ptr(self.b3), self.n_types, ctypes.c_float(m1),
                                      ctypes.c_float(m2), ctypes.c_float(m3), _ptr(self.allowed), _ptr(e_flat), cap,
                                      _ptr(valence), _ptr(stats), _stream())
        _check(eng.lib, rc, 'dndm_bond_orders')
        out = {'valence': valence, 'n_bonds': stats[:, 0], 'n_components': stats[:, 1], 'largest_component': stats[:, 2],
               'valence_violations': stats[:, 3]}
        if return_matrices:
            mats, off = [], 0
            for k in sizes.tolist():
                mats.append(e_flat[off:off + k * k].view(k, k))
                off += k * k
            out['E'] = mats
            out['E_flat'], out['sizes'] = e_flat[:cap], sizes
        return out

    def keep_mask(self, stats: Mapping[str, torch.Tensor], sizes: torch.Tensor, min_fragment: float = 0.5,
                  max_violations: int = 0) -> torch.Tensor:
        """A conservative candidate filter: keep molecules whose largest bonded fragment holds at least
        ``min_fragment`` of the atoms and with at most ``max_violations`` over-valent atoms."""
        frac = stats['largest_component'].float() / sizes.to(stats['largest_component'].device).clamp(min=1).float()
        return (frac >= min_fragment) & (stats['valence_violations'] <= max_violations)


class PrefilteredReward:
    """``reward_fn`` wrapper for the guided samplers: one ``dndm_bond_orders`` launch over all
    candidates of an SPSA / ATP round, and only the molecules that pass ``BondPerception.keep_mask`` are handed to the
    (expensive, host-side: OpenBabel + RDKit) scorer -- the rest get ``rejected_score``, the value the reference's
    ``my_reward_function`` assigns to molecules it cannot build (0).  Opt-in: it changes the result only if the scorer
    would have given a rejected molecule something else than ``rejected_score``.

    ``reward_fn(x_lig [n,3], atom_types [n], lig_mask [n]) -> list[float]`` as everywhere in ``sampler.py``."""

    def __init__(self, reward_fn, perception: BondPerception, rejected_score: float = 0.0, min_fragment: float = 0.5,
                 max_violations: int = 0):
        self.reward_fn = reward_fn
        self.perception = perception
        self.rejected_score = float(rejected_score)
        self.min_fragment = min_fragment
        self.max_violations = max_violations
        self.scored = 0
        self.rejected = 0

    @torch.no_grad()
    def __call__(self, x_lig: torch.Tensor, atom_types: torch.Tensor, lig_mask: torch.Tensor):
        n = int(lig_mask.max().item()) + 1 if lig_mask.numel() else 0
        if n == 0:
            return []
        x = x_lig[:, :3].contiguous().float()
        stats = self.perception(x, atom_types, lig_mask, n, return_matrices=False)
        sizes = torch.bincount(lig_mask, minlength=n)
        keep = self.perception.keep_mask(stats, sizes, self.min_fragment, self.max_violations)
        n_keep = int(keep.sum().item())
        self.scored += n_keep
        self.rejected += n - n_keep
        if n_keep == n:
            return list(self.reward_fn(x_lig, atom_types, lig_mask))
        out = [self.rejected_score] * n
        if n_keep:
            atom_keep = keep[lig_mask]
            new_id = torch.cumsum(keep.long(), 0) - 1
            scores = self.reward_fn(x_lig[atom_keep], atom_types[atom_keep], new_id[lig_mask][atom_keep])
            for i, s in zip(torch.nonzero(keep).flatten().tolist(), scores):
                out[i] = float(s)
        return out
