"""Pocket ingest and batching for the sampler: what sits between a PDB file and the first
denoiser call in the reference, without BioPython / RDKit.

Mirrors, with the same names and argument meaning:

* ``utils.get_pocket_from_ligand``                    utils.py:102-127
* ``utils.num_nodes_to_batch_mask`` / ``batch_to_list``  utils.py:130-153
* ``LigandPocketDDPM.prepare_pocket``                 lightning_modules.py:763-801
* the pocket selection of ``generate_ligands``        lightning_modules.py:846-858
* ``DistributionNodes`` (ligand-size prior)           equivariant_diffusion/en_diffusion.py:963-1033

and adds ``PocketCache``: the reference re-parses the PDB file and re-loads the checkpoint for every batch of one pocket
(`generate_ligands.py:98-105`, `my_test.py:68-90`); here a parsed pocket stays on the device, keyed by file + selection.

The static pocket--pocket part of the radius graph is deliberately NOT cached with the pocket: the reference recomputes
every distance from the translated coordinates of the current step (`dynamics.py:169-187`), and the fp32 rounding of
``x - com`` moves pairs that sit on the 5 A cutoff (PDB coordinates have 1e-3 A resolution, so exact ties exist) -- a cached
edge set would not be bit-exact with the reference's.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .datasets import THREE_TO_ONE

FLOAT_TYPE = torch.float32      # constants.py:6-7
INT_TYPE = torch.int64


@dataclass
class Atom:
    name: str
    element: str
    coord: np.ndarray           # float32 [3]
    occupancy: float = 1.0


@dataclass
class Residue:
    """One residue of the first model; ``id`` follows BioPython's ``(hetero flag, resseq, icode)``."""
    chain: str
    resname: str
    id: Tuple[str, int, str]
    atoms: List[Atom] = field(default_factory=list)

    def get_resname(self) -> str:
        return self.resname

    def get_atoms(self):
        return iter(self.atoms)

    def coords(self) -> np.ndarray:
        return np.stack([a.coord for a in self.atoms]).astype(np.float32)


def _guess_element(atom_name: str, is_het: bool) -> str:
    """Element from the atom-name columns 13-16 when columns 77-78 are blank (old PDB files): names that start in
    column 14 (or with a digit) carry a one-letter element; a letter in column 13 is a two-letter element for hetero
    atoms ("CA  " calcium, "BR  ") and the first letter otherwise ("HG21")."""
    letters = ''.join(c for c in atom_name if c.isalpha())
    if not letters:
        return 'X'
    if atom_name[:1].isalpha() and is_het:
        return letters[:2].capitalize()
    return letters[0].upper()


def parse_pdb(path: Union[str, os.PathLike]) -> List[Residue]:
    """ATOM / HETATM records of the first model, in file order (what ``PDBParser(QUIET=True).get_structure(..)[0]``
    iterates, lightning_modules.py:846).  Alternate locations collapse to the one with the highest occupancy (the first on
    ties), which is the atom BioPython's ``DisorderedAtom`` exposes."""
    residues: "OrderedDict[Tuple[str, Tuple[str, int, str]], Residue]" = OrderedDict()
    slots: Dict[Tuple, int] = {}
    with open(path) as f:
        for line in f:
            rec = line[:6]
            if rec == 'ENDMDL':
                break
            if rec not in ('ATOM  ', 'HETATM'):
                continue
            is_het = rec == 'HETATM'
            name = line[12:16]
            resname = line[17:20].strip()
            chain = line[21]
            resseq = int(line[22:26])
            icode = line[26] if len(line) > 26 else ' '
            het = ' '
            if is_het:
                het = 'W' if resname in ('HOH', 'WAT') else 'H_' + resname
            xyz = np.array([float(line[30:38]), float(line[38:46]), float(line[46:54])], np.float32)
            try:
                occ = float(line[54:60])
            except ValueError:
                occ = 1.0
            el = line[76:78].strip().capitalize() if len(line) >= 78 else ''
            if not el:
                el = _guess_element(name, is_het)
            key = (chain, (het, resseq, icode))
            res = residues.get(key)
            if res is None:
                res = residues[key] = Residue(chain, resname, (het, resseq, icode))
            akey = key + (name.strip(),)
            atom = Atom(name.strip(), el, xyz, occ)
            if akey in slots:                                   # alternate location of an atom already seen
                if occ > res.atoms[slots[akey]].occupancy:
                    res.atoms[slots[akey]] = atom
            else:
                slots[akey] = len(res.atoms)
                res.atoms.append(atom)
    return list(residues.values())


def read_sdf_coords(path: Union[str, os.PathLike]) -> np.ndarray:
    """Coordinates of the first molecule of a V2000 SDF file (``Chem.SDMolSupplier(..)[0].GetConformer()``)."""
    with open(path) as f:
        lines = f.read().splitlines()
    n = int(lines[3][:3])
    return np.array([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in lines[4:4 + n]], np.float32)


def is_aa(resname: str, standard: bool = True) -> bool:
    return resname.upper() in THREE_TO_ONE


def _min_dist(a: np.ndarray, b: np.ndarray) -> float:
    d = a[:, None, :].astype(np.float32) - b[None, :, :].astype(np.float32)
    return float(np.sqrt((d * d).sum(-1)).min())


def get_pocket_from_ligand(pdb_model: Sequence[Residue], ligand: str, dist_cutoff: float = 8.0) -> List[Residue]:
    """utils.py:102-127.  ``ligand``: path of an SDF file, or ``<chain>:<resi>`` of a residue inside the PDB file.  Returns
    the standard amino-acid residues with an atom closer than ``dist_cutoff`` to the ligand; as in the reference, every
    residue whose sequence number equals the ligand's is skipped, whatever its chain."""
    if str(ligand).endswith('.sdf'):
        ligand_coords = read_sdf_coords(ligand)
        resi = None
    else:
        chain, resi_s = str(ligand).split(':')
        resi = int(resi_s)
        hits = [r for r in pdb_model if r.chain == chain and r.id[1] == resi]
        assert len(hits) == 1                                   # get_residue_with_resi, utils.py:95-98
        ligand_coords = hits[0].coords()
    pocket = []
    for residue in pdb_model:
        if residue.id[1] == resi:
            continue
        if is_aa(residue.get_resname(), standard=True) and _min_dist(residue.coords(), ligand_coords) < dist_cutoff:
            pocket.append(residue)
    return pocket


def residues_from_ids(pdb_model: Sequence[Residue], pocket_ids: Sequence[str]) -> List[Residue]:
    """``pdb_struct[chain][(' ', resi, ' ')]`` for every ``<chain>:<resi>`` (lightning_modules.py:849-852)."""
    index = {(r.chain, r.id): r for r in pdb_model}
    out = []
    for x in pocket_ids:
        chain, resi = x.split(':')
        key = (chain, (' ', int(resi), ' '))
        if key not in index:
            raise KeyError(key)
        out.append(index[key])
    return out


def num_nodes_to_batch_mask(n_samples: int, num_nodes, device) -> torch.Tensor:
    """utils.py:145-153."""
    assert isinstance(num_nodes, int) or len(num_nodes) == n_samples
    if isinstance(num_nodes, torch.Tensor):
        num_nodes = num_nodes.to(device)
    sample_inds = torch.arange(n_samples, device=device)
    return torch.repeat_interleave(sample_inds, num_nodes)


def batch_to_list(data: torch.Tensor, batch_mask: torch.Tensor):
    """utils.py:130-142."""
    idx = torch.argsort(batch_mask, stable=True)
    batch_mask = batch_mask[idx]
    data = data[idx]
    chunk_sizes = torch.unique(batch_mask, return_counts=True)[1].tolist()
    return torch.split(data, chunk_sizes)


def pocket_arrays(residues: Sequence[Residue], pocket_type_encoder: Mapping[str, int],
                  pocket_representation: str = 'full-atom') -> Tuple[np.ndarray, np.ndarray]:
    """Coordinates [n,3] fp32 and vocabulary indices [n] of one pocket (the un-repeated half of ``prepare_pocket``)."""
    if pocket_representation == 'CA':
        xs, ts = [], []
        for res in residues:
            ca = [a for a in res.atoms if a.name == 'CA']
            if not ca:
                raise KeyError('CA')
            xs.append(ca[0].coord)
            ts.append(pocket_type_encoder[THREE_TO_ONE[res.get_resname().upper()]])
    else:
        atoms = [a for res in residues for a in res.get_atoms()
                 if (a.element.capitalize() in pocket_type_encoder or a.element != 'H')]
        xs = [a.coord for a in atoms]
        ts = [pocket_type_encoder[a.element.capitalize()] for a in atoms]     # KeyError for unknown heavy elements, as upstream
    return np.asarray(xs, np.float32).reshape(-1, 3), np.asarray(ts, np.int64)


def repeat_pocket(coord: torch.Tensor, types: torch.Tensor, n_types: int, repeats: int) -> Dict[str, torch.Tensor]:
    dev = coord.device
    one_hot = torch.nn.functional.one_hot(types, num_classes=n_types)
    n = int(coord.shape[0])
    return {
        'x': coord.repeat(repeats, 1),
        'one_hot': one_hot.repeat(repeats, 1),
        'size': torch.tensor([n] * repeats, device=dev, dtype=INT_TYPE),
        'mask': torch.repeat_interleave(torch.arange(repeats, device=dev, dtype=INT_TYPE), n),
    }


def prepare_pocket(biopython_residues: Sequence[Residue], repeats: int = 1, *, pocket_type_encoder: Mapping[str, int],
                   device='cpu', pocket_representation: str = 'full-atom') -> Dict[str, torch.Tensor]:
    """lightning_modules.py:763-801: dict with 'x' [repeats*n,3] fp32, 'one_hot' [repeats*n,n_types] int64,
    'size' [repeats] int64, 'mask' [repeats*n] int64."""
    x, t = pocket_arrays(biopython_residues, pocket_type_encoder, pocket_representation)
    return repeat_pocket(torch.tensor(x, device=device, dtype=FLOAT_TYPE), torch.tensor(t, device=device),
                         len(pocket_type_encoder), repeats)


class PocketCache:
    """Parsed pockets resident on the device, keyed by (file, mtime, selection).  ``get`` returns the ``prepare_pocket``
    dict for ``repeats`` copies; the parse, the residue selection and the host-to-device copy happen once per pocket."""

    def __init__(self, pocket_type_encoder: Mapping[str, int], device='cuda', pocket_representation: str = 'full-atom',
                 capacity: int = 256):
        self.encoder = dict(pocket_type_encoder)
        self.device = torch.device(device)
        self.representation = pocket_representation
        self.capacity = capacity
        self._entries: "OrderedDict[Tuple, Tuple[torch.Tensor, torch.Tensor]]" = OrderedDict()
        self.hits = 0
        self.misses = 0

    def _key(self, pdb_file, pocket_ids, ref_ligand):
        st = os.stat(pdb_file)
        sel = ('ids',) + tuple(pocket_ids) if pocket_ids is not None else ('ref', str(ref_ligand))
        if pocket_ids is None and str(ref_ligand).endswith('.sdf'):
            sel += (os.stat(ref_ligand).st_mtime_ns,)
        return (os.path.abspath(str(pdb_file)), st.st_mtime_ns, st.st_size) + sel

    def base(self, pdb_file, pocket_ids: Optional[Sequence[str]] = None, ref_ligand: Optional[str] = None):
        assert (pocket_ids is None) ^ (ref_ligand is None)      # lightning_modules.py:841
        key = self._key(pdb_file, pocket_ids, ref_ligand)
        hit = self._entries.get(key)
        if hit is not None:
            self.hits += 1
            self._entries.move_to_end(key)
            return hit
        self.misses += 1
        model = parse_pdb(pdb_file)
        residues = residues_from_ids(model, pocket_ids) if pocket_ids is not None else \
            get_pocket_from_ligand(model, ref_ligand)
        x, t = pocket_arrays(residues, self.encoder, self.representation)
        entry = (torch.from_numpy(x).to(self.device, FLOAT_TYPE), torch.from_numpy(t).to(self.device))
        self._entries[key] = entry
        while len(self._entries) > self.capacity:
            self._entries.popitem(last=False)
        return entry

    def get(self, pdb_file, pocket_ids: Optional[Sequence[str]] = None, ref_ligand: Optional[str] = None,
            repeats: int = 1) -> Dict[str, torch.Tensor]:
        coord, types = self.base(pdb_file, pocket_ids, ref_ligand)
        return repeat_pocket(coord, types, len(self.encoder), repeats)


class DistributionNodes:
    """Joint histogram prior over (ligand size, pocket size), en_diffusion.py:963-1033.  Sampling goes through the same
    sequence of ``torch.multinomial`` draws on the CPU generator as the reference's list of ``Categorical`` objects, so
    equal seeds give equal sizes."""

    def __init__(self, histogram):
        histogram = torch.as_tensor(np.asarray(histogram)).float()
        histogram = histogram + 1e-3
        prob = histogram / histogram.sum()
        n1, n2 = prob.shape
        self.idx_to_n_nodes = torch.stack(torch.meshgrid(torch.arange(n1), torch.arange(n2), indexing='ij'), -1).view(-1, 2)
        self.prob = prob
        # Categorical normalises its argument; keep the normalised rows / columns for sampling and log-probabilities
        self._flat = prob.view(-1) / prob.view(-1).sum()
        self._n1_given_n2 = (prob / prob.sum(0, keepdim=True)).T.contiguous()      # [n2, n1]
        self._n2_given_n1 = (prob / prob.sum(1, keepdim=True)).contiguous()        # [n1, n2]

    def entropy(self) -> float:
        p = self._flat
        return float(-(p * torch.log(p)).sum())

    def sample(self, n_samples: int = 1):
        idx = torch.multinomial(self._flat, n_samples, True)
        num_nodes_lig, num_nodes_pocket = self.idx_to_n_nodes[idx].T
        return num_nodes_lig, num_nodes_pocket

    def sample_conditional(self, n1=None, n2=None) -> torch.Tensor:
        assert (n1 is None) ^ (n2 is None), "Exactly one input argument must be None"
        table = self._n1_given_n2 if n2 is not None else self._n2_given_n1
        c = n2 if n2 is not None else n1
        draws = [int(torch.multinomial(table[int(i)], 1, True)) for i in c.tolist()]
        return torch.tensor(draws, device=c.device)

    def log_prob(self, batch_n_nodes_1: torch.Tensor, batch_n_nodes_2: torch.Tensor) -> torch.Tensor:
        assert batch_n_nodes_1.dim() == 1 and batch_n_nodes_2.dim() == 1
        idx = batch_n_nodes_1.cpu().long() * self.prob.shape[1] + batch_n_nodes_2.cpu().long()
        return torch.log(self._flat[idx]).to(batch_n_nodes_1.device)

    def log_prob_n1_given_n2(self, n1: torch.Tensor, n2: torch.Tensor) -> torch.Tensor:
        assert n1.dim() == 1 and n2.dim() == 1
        return torch.log(self._n1_given_n2[n2.cpu().long(), n1.cpu().long()]).to(n1.device)

    def log_prob_n2_given_n1(self, n2: torch.Tensor, n1: torch.Tensor) -> torch.Tensor:
        assert n1.dim() == 1 and n2.dim() == 1
        return torch.log(self._n2_given_n1[n1.cpu().long(), n2.cpu().long()]).to(n2.device)
