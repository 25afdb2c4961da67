"""``dataset_info`` of the crossdock full-atom model (reference ``constants.py:170-189``, ``dataset_params['crossdock']``).

Only the entries the hot path and its two neighbours (ingest, output) read: the atom / pocket vocabularies and the
typical bond lengths in pm from which ``get_bond_order_batch`` (analysis/molecule_builder.py:30-55) decides bond orders.
The tables are stored once per unordered element pair and expanded to the symmetric ``[n_types, n_types]`` matrices the
reference indexes (`tests/test_ingest_output.py` checks them against the reference-generated ``tests/golden/bonds.npz``).
"""
from __future__ import annotations

from typing import Dict, List

ATOM_DECODER: List[str] = ['C', 'N', 'O', 'S', 'B', 'Br', 'Cl', 'P', 'I', 'F']
AA_DECODER: List[str] = ['A', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'K', 'L', 'M', 'N', 'P', 'Q', 'R', 'S', 'T', 'V', 'W', 'Y']
THREE_TO_ONE: Dict[str, str] = {
    'ALA': 'A', 'CYS': 'C', 'ASP': 'D', 'GLU': 'E', 'PHE': 'F', 'GLY': 'G', 'HIS': 'H', 'ILE': 'I', 'LYS': 'K', 'LEU': 'L',
    'MET': 'M', 'ASN': 'N', 'PRO': 'P', 'GLN': 'Q', 'ARG': 'R', 'SER': 'S', 'THR': 'T', 'VAL': 'V', 'TRP': 'W', 'TYR': 'Y'}

# typical bond lengths in pm, one entry per unordered pair
_SINGLE = {
    'C': {'C': 154, 'N': 147, 'O': 143, 'S': 182, 'Br': 194, 'Cl': 177, 'P': 184, 'I': 214, 'F': 135},
    'N': {'N': 145, 'O': 140, 'S': 168, 'Br': 214, 'Cl': 175, 'P': 177, 'I': 222, 'F': 136},
    'O': {'O': 148, 'S': 151, 'Br': 172, 'Cl': 164, 'P': 163, 'I': 194, 'F': 142},
    'S': {'S': 204, 'Br': 225, 'Cl': 207, 'P': 210, 'I': 234, 'F': 158},
    'B': {'Cl': 175},
    'Br': {'Br': 228, 'Cl': 214, 'P': 222, 'F': 178},
    'Cl': {'Cl': 199, 'P': 203, 'F': 166},
    'P': {'P': 221, 'F': 156},
    'I': {'I': 266, 'F': 187},
    'F': {'F': 142},
}
_DOUBLE = {
    'C': {'C': 134, 'N': 129, 'O': 120, 'S': 160},
    'N': {'N': 125, 'O': 121},
    'O': {'O': 121, 'P': 150},
    'S': {'P': 186},
}
_TRIPLE = {
    'C': {'C': 120, 'N': 116, 'O': 113},
    'N': {'N': 110},
}


def _matrix(pairs: Dict[str, Dict[str, int]], decoder: List[str]) -> List[List[float]]:
    idx = {a: i for i, a in enumerate(decoder)}
    m = [[0.0] * len(decoder) for _ in decoder]
    for a, row in pairs.items():
        for b, v in row.items():
            m[idx[a]][idx[b]] = m[idx[b]][idx[a]] = float(v)
    return m


def crossdock_dataset_info(pocket_representation: str = 'full-atom') -> Dict[str, object]:
    """The dict the reference passes around as ``dataset_info`` (``lightning_modules.py:76``).  In full-atom mode the
    pocket uses the atom vocabulary (``lightning_modules.py:93-98``), in CA mode the one-letter residue codes."""
    info = {
        'atom_encoder': {a: i for i, a in enumerate(ATOM_DECODER)},
        'atom_decoder': list(ATOM_DECODER),
        'aa_encoder': {a: i for i, a in enumerate(AA_DECODER)},
        'aa_decoder': list(AA_DECODER),
        'bonds1': _matrix(_SINGLE, ATOM_DECODER),
        'bonds2': _matrix(_DOUBLE, ATOM_DECODER),
        'bonds3': _matrix(_TRIPLE, ATOM_DECODER),
    }
    info['pocket_encoder'] = info['atom_encoder'] if pocket_representation != 'CA' else info['aa_encoder']
    return info
