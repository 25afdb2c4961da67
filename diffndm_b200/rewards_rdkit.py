"""Host chemistry of the guided samplers: the reference's molecule hand-off and rewards as a picklable ``score_one``.

Restates, with RDKit / OpenBabel imported lazily inside the worker -- the GPU image has neither,
the module imports without them and raises a clear error only when a molecule is actually scored:

* ``handle_to_mol``        conditional_model.py:845-882   translate back to the input frame, ``build_molecule`` per
                                                          ligand (OpenBabel bond perception, analysis/molecule_builder.py:58-97,
                                                          139-159), ``process_molecule`` (:160-209)
* ``my_reward_for_SPSA``   conditional_model.py:816-843   2 QED + 3 SA + Lipinski / 5
* ``my_reward_for_SVDD``   conditional_model.py:622-653   2 QED + 2 SA + logP window + Lipinski / 5
* ``MoleculeProperties``   analysis/metrics.py:135-178, 282-368 (``evaluate_new``: a molecule that fails sanitisation scores 0
                                                          on every metric)

The translation of ``handle_to_mol`` does not change any of these scores (they are functions of the molecular graph), so
``score_one`` takes the sampler-frame coordinates; ``RdkitReward`` keeps ``pocket_com_before`` only to hand back molecules
in the input frame (``to_mols``).  Differences from the reference, on purpose: OpenBabel is fed the xyz block from memory
instead of through a temporary file, and molecules are scored by ``hostpool.PooledReward`` workers in parallel.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from .datasets import ATOM_DECODER


def _require():
    try:
        from rdkit import Chem  # noqa: F401
    except Exception as e:                                      # pragma: no cover - depends on the host image
        raise RuntimeError('diffndm_b200.rewards_rdkit needs RDKit (and OpenBabel for bond perception) on the host; '
                           'pass another reward_fn to the sampler, e.g. hostpool.PooledReward(your_score_one)') from e


def rdkit_available() -> bool:
    try:
        import rdkit  # noqa: F401
        return True
    except Exception:
        return False


def xyz_block(positions: np.ndarray, atom_types: Sequence[int], decoder: Sequence[str] = ATOM_DECODER) -> str:
    """utils.write_xyz_file (utils.py:63-70) as a string."""
    lines = [str(len(positions)), '']
    for t, p in zip(atom_types, positions):
        lines.append(f'{decoder[int(t)]} {float(p[0]):.3f} {float(p[1]):.3f} {float(p[2]):.3f}')
    return '\n'.join(lines) + '\n'


def build_molecule(positions: np.ndarray, atom_types: Sequence[int], decoder: Sequence[str] = ATOM_DECODER):
    """make_mol_openbabel (molecule_builder.py:58-97): OpenBabel perceives the bonds of the xyz block, the molecule is
    rebuilt atom by atom in RDKit (which drops OpenBabel's radicals) with the conformer attached."""
    _require()
    from rdkit import Chem
    try:
        from openbabel import openbabel
    except Exception:                                           # pragma: no cover
        import openbabel
    conv = openbabel.OBConversion()
    conv.SetInAndOutFormats('xyz', 'sdf')
    ob = openbabel.OBMol()
    conv.ReadString(ob, xyz_block(positions, atom_types, decoder))
    tmp = Chem.MolFromMolBlock(conv.WriteString(ob), sanitize=False)
    mol = Chem.RWMol()
    for atom in tmp.GetAtoms():
        mol.AddAtom(Chem.Atom(atom.GetSymbol()))
    mol.AddConformer(tmp.GetConformer(0))
    for bond in tmp.GetBonds():
        mol.AddBond(bond.GetBeginAtomIdx(), bond.GetEndAtomIdx(), bond.GetBondType())
    return mol


def process_molecule(rdmol, sanitize: bool = False, relax_iter: int = 0, largest_frag: bool = False):
    """molecule_builder.py:160-209 (add_hydrogens=False as in handle_to_mol): a copy of the molecule, optionally
    sanitised, reduced to its largest fragment and UFF-relaxed; None when a filter fails."""
    _require()
    from rdkit import Chem
    from rdkit.Chem.rdForceFieldHelpers import UFFHasAllMoleculeParams, UFFOptimizeMolecule
    mol = Chem.Mol(rdmol)
    if sanitize:
        try:
            Chem.SanitizeMol(mol)
        except ValueError:
            return None
    if largest_frag:
        frags = Chem.GetMolFrags(mol, asMols=True, sanitizeFrags=False)
        mol = max(frags, default=mol, key=lambda m: m.GetNumAtoms())
        if sanitize:
            try:
                Chem.SanitizeMol(mol)
            except ValueError:
                return None
    if relax_iter > 0:
        if not UFFHasAllMoleculeParams(mol):
            return None
        try:
            UFFOptimizeMolecule(mol, maxIters=relax_iter)
            if sanitize:
                Chem.SanitizeMol(mol)
        except (RuntimeError, ValueError):
            return None
    return mol


def molecule_properties(mol):
    """(QED, SA, logP, Lipinski) of MoleculeProperties.evaluate_new (metrics.py:282-368): zeros when sanitisation fails.
    SA = round((10 - sascorer) / 9, 2) (:148-150); the Lipinski count reproduces the reference's fourth rule as written
    (``(logp := MolLogP >= -2) & (logp <= 5)`` binds the comparison, so only the lower bound is tested, :171)."""
    _require()
    from rdkit import Chem
    from rdkit.Chem import Crippen, Descriptors, Lipinski, QED, rdMolDescriptors
    try:
        from rdkit.Contrib.SA_Score import sascorer
    except Exception:                                           # pragma: no cover
        import sascorer
    try:
        Chem.SanitizeMol(mol)
    except Exception:
        return 0.0, 0.0, 0.0, 0.0
    logp = Crippen.MolLogP(mol)
    rules = [Descriptors.ExactMolWt(mol) < 500, Lipinski.NumHDonors(mol) <= 5, Lipinski.NumHAcceptors(mol) <= 10,
             logp >= -2, rdMolDescriptors.CalcNumRotatableBonds(mol) <= 10]
    return float(QED.qed(mol)), round((10 - sascorer.calculateScore(mol)) / 9, 2), float(logp), float(sum(int(r) for r in rules))


def reward_spsa(props) -> float:
    qed, sa, _, lip = props
    return qed * 2 + sa * 3 + lip / 5                           # conditional_model.py:832-840


def reward_svdd(props) -> float:
    qed, sa, logp, lip = props
    sg = lambda v: 1.0 / (1.0 + math.exp(-v))
    k = 20
    return qed * 2 + sa * 2 + sg(k * (logp + 1)) * sg(-k * (logp - 5)) + lip / 5        # :643-651


class ScoreOne:
    """Picklable ``score_one(positions, atom_types) -> float`` for ``hostpool.PooledReward``."""

    def __init__(self, kind: str = 'spsa', sanitize: bool = False, relax_iter: int = 0, largest_frag: bool = False,
                 decoder: Sequence[str] = ATOM_DECODER):
        assert kind in ('spsa', 'svdd')
        self.kind, self.sanitize, self.relax_iter, self.largest_frag, self.decoder = kind, sanitize, relax_iter, largest_frag, list(decoder)

    def __call__(self, positions: np.ndarray, atom_types: np.ndarray) -> float:
        mol = process_molecule(build_molecule(positions, atom_types, self.decoder), sanitize=self.sanitize,
                               relax_iter=self.relax_iter, largest_frag=self.largest_frag)
        if mol is None:          # the reference drops the molecule and its reward list gets shorter (a latent bug); score 0
            return 0.0
        props = molecule_properties(mol)
        return reward_spsa(props) if self.kind == 'spsa' else reward_svdd(props)


def make_reward_fn(kind: str, sanitize=False, relax_iter=0, largest_frag=False, workers: int = 8):
    """``reward_fn(x_lig, atom_types, lig_mask)`` for ``ConditionalSampler`` scoring with the reference's chemistry on a
    process pool (``submit`` is offered, so scoring overlaps GPU denoising)."""
    from .hostpool import PooledReward
    return PooledReward(ScoreOne(kind, sanitize, relax_iter, largest_frag), workers=workers)


class GuidanceRewards:
    """The pair the reference uses: SPSA updates score with ``my_reward_for_SPSA``, ATP selections with
    ``my_reward_for_SVDD``.  The sampler asks ``for_event(kind)``."""

    def __init__(self, sanitize=False, relax_iter=0, largest_frag=False, workers: int = 8):
        self.args = (sanitize, relax_iter, largest_frag)
        self.workers = workers
        self._fns = {}

    def for_event(self, kind: str):
        if kind not in self._fns:
            self._fns[kind] = make_reward_fn(kind, *self.args, workers=self.workers)
        return self._fns[kind]

    def close(self):
        for f in self._fns.values():
            f.close()
        self._fns = {}
