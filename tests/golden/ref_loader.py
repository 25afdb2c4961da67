"""Import the UNMODIFIED reference (``/root/reference``) behind shims.  Build-container only.

The reference needs ``torch_scatter``, ``rdkit``, ``openbabel``, ``Bio``, ``pytorch_lightning`` which are
not installed here.  The numerical path only uses ``torch_scatter.scatter_add/scatter_mean``,
which are restated below as plain segment sum / mean; the chemistry packages are mocked because the denoiser
never calls them.  Nothing in this file travels to the GPU box (``/root/reference`` is absent there); the
fixtures it produces do.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get('DNDM_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'equivariant_diffusion'))


def _scatter_add(src, index, dim=0, dim_size=None):
    assert dim == 0
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def _scatter_mean(src, index, dim=0, dim_size=None):
    s = _scatter_add(src, index, dim, dim_size)
    cnt = torch.bincount(index, minlength=s.shape[0]).clamp(min=1).to(src.dtype)
    return s / cnt.view((-1,) + (1,) * (src.dim() - 1))


def install_shims():
    ts = types.ModuleType('torch_scatter')
    ts.scatter_add = _scatter_add
    ts.scatter_mean = _scatter_mean
    sys.modules.setdefault('torch_scatter', ts)
    for name in ['rdkit', 'rdkit.Chem', 'rdkit.Chem.rdMolTransforms', 'rdkit.Chem.Descriptors', 'rdkit.Chem.Crippen',
                 'rdkit.Chem.Lipinski', 'rdkit.Chem.QED', 'rdkit.Chem.AllChem', 'rdkit.Chem.rdMolDescriptors',
                 'rdkit.Chem.rdchem', 'rdkit.DataStructs', 'rdkit.six', 'rdkit.six.moves', 'rdkit.Chem.rdForceFieldHelpers',
                 'rdkit.Chem.Draw', 'openbabel', 'Bio', 'Bio.PDB', 'Bio.PDB.Polypeptide', 'Bio.PDB.PDBParser',
                 'networkx', 'networkx.algorithms', 'networkx.algorithms.isomorphism', 'imageio', 'wandb',
                 'matplotlib', 'matplotlib.pyplot', 'seaborn']:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference():
    """Returns (EGNNDynamics, ConditionalDDPM) classes of the reference."""
    install_shims()
    with contextlib.redirect_stdout(io.StringIO()):
        from equivariant_diffusion.dynamics import EGNNDynamics
        from equivariant_diffusion.conditional_model import ConditionalDDPM
    return EGNNDynamics, ConditionalDDPM


def build_reference_model(cfg, weights, dtype=torch.float32, timesteps=500, untie_heads=False):
    """EGNNDynamics + ConditionalDDPM with the kwargs of lightning_modules.py:138-174 and the given weights.

    The reference builds ONE ``nn.Linear(hidden_nf, 1, bias=False)`` and puts it at the end of both ``coord_mlp`` and
    ``cross_product_mlp`` (egnn_new.py:78-92): the two state-dict keys alias one Parameter, and ``load_state_dict`` leaves
    both heads with whichever key it copied last.  ``untie_heads`` gives ``cross_product_mlp`` its own Linear object before
    loading, so that a table whose two keys differ is represented as written (the engine packs the keys separately)."""
    EGNNDynamics, ConditionalDDPM = load_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        dyn = EGNNDynamics(
            atom_nf=cfg.atom_nf, residue_nf=cfg.residue_nf, n_dims=3, joint_nf=cfg.joint_nf, device='cpu',
            hidden_nf=cfg.hidden_nf, act_fn=torch.nn.SiLU(), n_layers=cfg.n_layers, attention=True, tanh=True,
            norm_constant=cfg.norm_constant, inv_sublayers=1, sin_embedding=False,
            normalization_factor=cfg.normalization_factor, aggregation_method='sum',
            edge_cutoff_ligand=cfg.edge_cutoff_ligand, edge_cutoff_pocket=cfg.edge_cutoff_pocket,
            edge_cutoff_interaction=cfg.edge_cutoff_interaction, update_pocket_coords=False,
            reflection_equivariant=False, edge_embedding_dim=cfg.edge_embedding_dim)
        if untie_heads:
            for i in range(cfg.n_layers):
                eq = dyn.egnn._modules[f'e_block_{i}']._modules['gcl_equiv']
                eq.cross_product_mlp[4] = torch.nn.Linear(cfg.hidden_nf, 1, bias=False)
        sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
        missing, unexpected = dyn.load_state_dict(sd, strict=True), None
        ddpm = ConditionalDDPM(dynamics=dyn, atom_nf=cfg.atom_nf, residue_nf=cfg.residue_nf, n_dims=3,
                               timesteps=timesteps, noise_schedule='polynomial_2', noise_precision=5.0e-4,
                               loss_type='l2', norm_values=[1, 4], size_histogram=np.ones((4, 4)))
    dyn.eval()
    ddpm.eval()
    if dtype == torch.float64:
        dyn.double()
    return dyn, ddpm
