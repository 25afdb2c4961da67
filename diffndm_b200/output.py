"""Molecule assembly and the SDF wire format, without RDKit.

Mirrors the tail of ``LigandPocketDDPM.generate_ligands`` (lightning_modules.py:928-949):

* ``build_molecules``   -- ``build_molecule(.., use_openbabel=False)`` = ``make_mol_edm`` (analysis/molecule_builder.py:100-136)
  for a whole batch: ONE ``dndm_bond_orders`` launch perceives the bonds of every molecule on the GPU, the host only
  collects the non-zero entries of the directed lower-triangular matrices (``torch.nonzero(A)`` order: row-major).
* ``process_molecule``  -- the ``largest_frag`` filter of analysis/molecule_builder.py:160-209.  ``sanitize``,
  ``add_hydrogens`` and ``relax_iter`` are RDKit / UFF operations and stay external (``NotImplementedError`` here; pass a
  ``mol_builder`` to ``LigandGenerator`` on a host that has RDKit to get the reference's exact objects).
* ``write_sdf_file`` / ``write_xyz_file`` -- utils.py:63-82.  V2000 molfile records (counts line, atom block with 4-decimal
  coordinates, bond block, ``M  END``, ``$$$$``), readable by RDKit's ``SDMolSupplier`` and by ``read_sdf``.

The reference's default bond perception is OpenBabel's (``make_mol_openbabel``); that is third-party chemistry outside the
repository and stays on the host exactly as in the reference.  The EDM distance rules are the reference's own second path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Union

import numpy as np
import torch

from .chem import BondPerception


@dataclass
class Molecule:
    """What the sampler's output needs from an ``rdkit.Chem.Mol``: atoms, one conformer, bonds."""
    symbols: List[str]
    positions: np.ndarray                      # [n,3] float32, Angstrom
    bonds: np.ndarray                          # [m,3] int: (begin, end, order) with begin > end (directed, tril)
    atom_types: Optional[np.ndarray] = None    # decoder indices
    props: Dict[str, str] = field(default_factory=dict)

    def GetNumAtoms(self) -> int:
        return len(self.symbols)

    def GetNumBonds(self) -> int:
        return int(self.bonds.shape[0])

    def fragment_labels(self) -> np.ndarray:
        """Connected components of the bond graph (``Chem.GetMolFrags``): label = smallest atom index of the fragment."""
        n = self.GetNumAtoms()
        parent = np.arange(n)

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a
        for i, j, _ in self.bonds:
            ri, rj = find(int(i)), find(int(j))
            if ri != rj:
                parent[max(ri, rj)] = min(ri, rj)
        return np.array([find(a) for a in range(n)], np.int64)

    def subset(self, keep: np.ndarray) -> 'Molecule':
        keep = np.asarray(keep, bool)
        new_index = np.cumsum(keep) - 1
        sel = [b for b in self.bonds if keep[b[0]] and keep[b[1]]]
        bonds = np.array([[new_index[i], new_index[j], o] for i, j, o in sel], np.int64).reshape(-1, 3)
        return Molecule([s for s, k in zip(self.symbols, keep) if k], self.positions[keep], bonds,
                        None if self.atom_types is None else self.atom_types[keep], dict(self.props))


def molecule_from_bond_matrix(positions: np.ndarray, atom_types: np.ndarray, E: np.ndarray, atom_decoder: Sequence[str]) -> Molecule:
    """``make_mol_edm`` after the bond-order matrix: atoms in order, bonds = ``nonzero(tril(E, -1))`` row-major."""
    E = np.tril(np.asarray(E), -1)
    ii, jj = np.nonzero(E)
    bonds = np.stack([ii, jj, E[ii, jj]], 1).astype(np.int64) if len(ii) else np.zeros((0, 3), np.int64)
    return Molecule([atom_decoder[int(t)] for t in atom_types], np.asarray(positions, np.float32), bonds,
                    np.asarray(atom_types, np.int64))


@torch.no_grad()
def build_molecules(x: torch.Tensor, atom_types: torch.Tensor, mol_mask: torch.Tensor, n_mols: int,
                    dataset_info: Mapping[str, object], perception: BondPerception) -> List[Molecule]:
    """Batch version of ``build_molecule(*mol_pc, dataset_info, add_coords=True, use_openbabel=False)``
    (lightning_modules.py:937-941).  x [N,3] fp32 CUDA, atom_types [N] int64, mol_mask [N] sorted int64."""
    out = perception(x, atom_types, mol_mask, n_mols, return_matrices=True)
    if n_mols and int(out['sizes'].max()) > 256:                # BOND_MAX_ATOMS of the kernel: such a block is left unwritten
        raise ValueError('build_molecules: a molecule has more than 256 atoms')
    xs = x.detach().cpu().numpy()
    ts = atom_types.detach().cpu().numpy()
    e_flat = out['E_flat'].cpu().numpy()                       # one device-to-host copy for the whole batch
    mols, off, eoff = [], 0, 0
    for k in out['sizes'].cpu().tolist():
        E = e_flat[eoff:eoff + k * k].reshape(k, k)
        mols.append(molecule_from_bond_matrix(xs[off:off + k, :3], ts[off:off + k], E, dataset_info['atom_decoder']))
        off += k
        eoff += k * k
    return mols


def process_molecule(mol: Molecule, add_hydrogens: bool = False, sanitize: bool = False, relax_iter: int = 0,
                     largest_frag: bool = False) -> Optional[Molecule]:
    """analysis/molecule_builder.py:160-209 restricted to what needs no RDKit: a copy, optionally reduced to its largest
    fragment (the first one on ties, like ``max(mol_frags, key=GetNumAtoms)`` over fragments in atom order)."""
    if sanitize or add_hydrogens or relax_iter > 0:
        raise NotImplementedError('sanitize / add_hydrogens / relax_iter are RDKit operations and stay host-side '
                                  '(external); pass mol_builder= to LigandGenerator on a host that has RDKit')
    if largest_frag and mol.GetNumAtoms() > 0:
        labels = mol.fragment_labels()
        uniq, counts = np.unique(labels, return_counts=True)        # sorted by smallest atom index = GetMolFrags order
        return mol.subset(labels == uniq[int(np.argmax(counts))])
    return mol.subset(np.ones(mol.GetNumAtoms(), bool))


def _mol_block(mol: Molecule, name: str = '') -> str:
    n, m = mol.GetNumAtoms(), mol.GetNumBonds()
    if n > 999 or m > 999:
        raise ValueError('V2000 molfiles hold at most 999 atoms / bonds')
    lines = [name, '     diffndm_b200          3D', '',
             f'{n:3d}{m:3d}  0  0  0  0  0  0  0  0999 V2000']
    for s, p in zip(mol.symbols, mol.positions):
        lines.append(f'{p[0]:10.4f}{p[1]:10.4f}{p[2]:10.4f} {s:<3s} 0  0  0  0  0  0  0  0  0  0  0  0')
    for i, j, o in mol.bonds:
        lines.append(f'{int(i) + 1:3d}{int(j) + 1:3d}{int(o):3d}  0')
    lines.append('M  END')
    for k, v in mol.props.items():
        lines += [f'>  <{k}>', str(v), '']
    lines.append('$$$$')
    return '\n'.join(lines) + '\n'


def write_sdf_file(sdf_path: Union[str, os.PathLike], molecules: Iterable[Optional[Molecule]]) -> int:
    """utils.py:72-82: one record per molecule that is not None, bond orders written as perceived (no kekulisation).
    Returns the number of records."""
    n = 0
    with open(sdf_path, 'w') as f:
        for mol in molecules:
            if mol is not None:
                f.write(_mol_block(mol))
                n += 1
    return n


def write_xyz_file(coords, atom_types: Sequence[str], filename: Union[str, os.PathLike]) -> None:
    """utils.py:63-70."""
    assert len(coords) == len(atom_types)
    out = f"{len(coords)}\n\n"
    for i in range(len(coords)):
        out += f"{atom_types[i]} {float(coords[i][0]):.3f} {float(coords[i][1]):.3f} {float(coords[i][2]):.3f}\n"
    with open(filename, 'w') as f:
        f.write(out)


def read_sdf(path: Union[str, os.PathLike]) -> List[Molecule]:
    """All V2000 records of an SDF file (the inverse of ``write_sdf_file``; also reads the reference's
    ``example/*.sdf`` and ``my_example_*`` outputs for distribution-level comparisons)."""
    with open(path) as f:
        text = f.read()
    mols = []
    for rec in text.split('$$$$'):
        lines = rec.strip('\n').split('\n')
        if len(lines) < 4 or 'V2000' not in rec:
            continue
        start = next(i for i, l in enumerate(lines) if 'V2000' in l)
        n, m = int(lines[start][:3]), int(lines[start][3:6])
        atoms = lines[start + 1:start + 1 + n]
        pos = np.array([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in atoms], np.float32).reshape(-1, 3)
        sym = [l[31:34].strip() for l in atoms]
        bonds = []
        for l in lines[start + 1 + n:start + 1 + n + m]:
            a, b, o = int(l[0:3]) - 1, int(l[3:6]) - 1, int(l[6:9])
            bonds.append((max(a, b), min(a, b), o))
        mols.append(Molecule(sym, pos, np.array(bonds, np.int64).reshape(-1, 3)))
    return mols
