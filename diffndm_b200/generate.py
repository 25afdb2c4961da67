"""``generate_ligands`` operator surface on top of the engine: PDB file in, molecules out.

Mirrors ``LigandPocketDDPM.generate_ligands`` (lightning_modules.py:803-949) and the batch loop of the
``generate_ligands.py`` script (:92-108) -- same argument names and meaning -- with the pocket served from a
``PocketCache`` (one parse + one upload per pocket instead of one per batch), the denoising loop on the B200 engine
(``ConditionalSampler``) and the bond perception of the whole batch in one GPU launch.  Host chemistry (OpenBabel bond
perception, RDKit sanitisation / UFF relaxation / QED-SA rewards) stays external: ``mol_builder`` and ``reward_fn`` are the
hooks a host that has those packages plugs them into.
"""
from __future__ import annotations

from typing import Callable, List, Mapping, Optional, Sequence

import torch

from . import ingest, output
from .chem import BondPerception
from .sampler import ConditionalSampler


def state_dict_from_checkpoint(ckpt, prefix: str = 'ddpm.dynamics.'):
    """Denoiser weights out of a Lightning checkpoint of the reference (``LigandPocketDDPM.load_from_checkpoint``,
    generate_ligands.py:57-58): ``ckpt`` is a path or the loaded dict; keys ``ddpm.dynamics.*`` lose their prefix
    (SURVEY.md section 9.1).  Returns (state, hyper_parameters or {})."""
    if not isinstance(ckpt, Mapping):
        ckpt = torch.load(ckpt, map_location='cpu', weights_only=False)
    sd = ckpt.get('state_dict', ckpt)
    state = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    if not state:
        raise KeyError(f'no "{prefix}*" entries in the checkpoint')
    return state, dict(ckpt.get('hyper_parameters', {}) or {})


def config_from_checkpoint(state: Mapping[str, object], hparams: Optional[Mapping[str, object]] = None):
    """``(DynamicsConfig, pocket_representation, sampler_kwargs)`` of a reference checkpoint: the sizes are read
    off the weight shapes (they cannot disagree with the tensors that will be loaded), cutoffs / normalisation / schedule from
    ``hyper_parameters['egnn_params' / 'diffusion_params']`` (``LigandPocketDDPM.__init__``, lightning_modules.py:32-57,
    138-174) where present -- a Namespace or a dict -- and from the crossdock_fullatom_cond defaults otherwise.  Covers every
    pocket-conditional configuration under ``configs/``; a joint model (``mode != 'pocket_conditioning'``) is refused."""
    from .weights import DynamicsConfig
    hp = dict(hparams or {})
    if hp.get('mode', 'pocket_conditioning') != 'pocket_conditioning':
        raise NotImplementedError(f"mode {hp['mode']!r}: only pocket-conditional checkpoints are on the engine's path")

    def shape(key):
        v = state[key]
        return tuple(v.shape)

    def field(group, name, default):
        g = hp.get(group)
        if g is None:
            return default
        v = g.get(name, default) if isinstance(g, Mapping) else getattr(g, name, default)
        return default if v is None and name not in ('edge_cutoff_ligand', 'edge_cutoff_pocket', 'edge_cutoff_interaction') else v

    hidden, d_in = shape('egnn.embedding.weight')
    n_layers = 1 + max(int(k.split('.')[1].split('_')[-1]) for k in state if k.startswith('egnn.e_block_'))
    base = DynamicsConfig()
    cfg = DynamicsConfig(
        atom_nf=shape('atom_encoder.0.weight')[1], residue_nf=shape('residue_encoder.0.weight')[1], joint_nf=d_in - 1,
        hidden_nf=hidden, n_layers=n_layers,
        edge_embedding_dim=(shape('edge_embedding.weight')[1] if 'edge_embedding.weight' in state else None),
        edge_cutoff_ligand=field('egnn_params', 'edge_cutoff_ligand', base.edge_cutoff_ligand),
        edge_cutoff_pocket=field('egnn_params', 'edge_cutoff_pocket', base.edge_cutoff_pocket),
        edge_cutoff_interaction=field('egnn_params', 'edge_cutoff_interaction', base.edge_cutoff_interaction),
        norm_constant=float(field('egnn_params', 'norm_constant', base.norm_constant)),
        normalization_factor=float(field('egnn_params', 'normalization_factor', base.normalization_factor)),
        attention=bool(field('egnn_params', 'attention', True)), tanh=bool(field('egnn_params', 'tanh', True)),
        reflection_equivariant=bool(field('egnn_params', 'reflection_equivariant', False)),
        inv_sublayers=int(field('egnn_params', 'inv_sublayers', 1)))
    rep = hp.get('pocket_representation') or ('CA' if cfg.residue_nf == 20 and cfg.atom_nf != 20 else 'full-atom')
    sampler_kwargs = dict(timesteps=int(field('diffusion_params', 'diffusion_steps', 500)),
                          noise_schedule=str(field('diffusion_params', 'diffusion_noise_schedule', 'polynomial_2')),
                          noise_precision=float(field('diffusion_params', 'diffusion_noise_precision', 5.0e-4)),
                          norm_values=tuple(field('diffusion_params', 'normalize_factors', (1.0, 4.0))))
    return cfg, rep, sampler_kwargs


class LigandGenerator:
    """The part of ``LigandPocketDDPM`` that ``generate_ligands.py`` / ``my_test.py`` / ``inpaint.py`` use at inference."""

    def __init__(self, sampler: ConditionalSampler, dataset_info: Mapping[str, object],
                 size_histogram=None, pocket_representation: str = 'full-atom',
                 pocket_cache: Optional[ingest.PocketCache] = None,
                 mol_builder: Optional[Callable] = None):
        self.ddpm = sampler
        self.dataset_info = dataset_info
        self.device = sampler.device
        self.x_dims = 3
        self.atom_nf = sampler.atom_nf
        self.pocket_representation = pocket_representation
        self.pocket_type_encoder = dataset_info.get('pocket_encoder') or (
            dataset_info['aa_encoder'] if pocket_representation == 'CA' else dataset_info['atom_encoder'])
        self.size_distribution = None if size_histogram is None else ingest.DistributionNodes(size_histogram)
        self.pockets = pocket_cache or ingest.PocketCache(self.pocket_type_encoder, self.device, pocket_representation)
        self.perception = BondPerception(sampler.engine, dataset_info)
        self.mol_builder = mol_builder

    @classmethod
    def from_checkpoint(cls, ckpt, dataset_info: Optional[Mapping[str, object]] = None, **kwargs) -> 'LigandGenerator':
        """``LigandPocketDDPM.load_from_checkpoint`` (generate_ligands.py:57-58) for the engine: weights, sizes, cutoffs,
        schedule and the ligand-size histogram all come out of the checkpoint (``config_from_checkpoint``).  ``dataset_info``
        defaults to the crossdock / bindingmoad table -- the two share the vocabularies and bond tables the path reads
        (constants.py:96-131, 169-187)."""
        from .datasets import crossdock_dataset_info
        from .engine import B200EGNNDynamics
        state, hparams = state_dict_from_checkpoint(ckpt)
        cfg, rep, sampler_kwargs = config_from_checkpoint(state, hparams)
        dyn = B200EGNNDynamics(cfg, state).eval()
        return cls(ConditionalSampler(dyn, **sampler_kwargs), dataset_info or crossdock_dataset_info(rep),
                   size_histogram=hparams.get('node_histogram'), pocket_representation=rep, **kwargs)

    # lightning_modules.py:763-801 (kept as a method because callers use it as one)
    def prepare_pocket(self, biopython_residues, repeats: int = 1):
        return ingest.prepare_pocket(biopython_residues, repeats, pocket_type_encoder=self.pocket_type_encoder,
                                     device=self.device, pocket_representation=self.pocket_representation)

    def _com(self, x: torch.Tensor, mask: torch.Tensor, n: int) -> torch.Tensor:
        from .sampler import segment_mean_sorted
        return segment_mean_sorted(x, mask, n)

    @torch.no_grad()
    def generate_ligands(self, pdb_file, n_samples: int, pocket_ids: Optional[Sequence[str]] = None,
                         ref_ligand: Optional[str] = None, num_nodes_lig=None, sanitize: bool = False,
                         largest_frag: bool = False, relax_iter: int = 0, timesteps: Optional[int] = None,
                         n_nodes_bias: int = 0, n_nodes_min: int = 0, svdd: int = 0, spsa: int = 0,
                         reward_fn: Optional[Callable] = None, return_tensors: bool = False, **kwargs) -> List:
        """Generate ligands given a pocket (lightning_modules.py:803-949).  ``pdb_file``, ``pocket_ids`` (list of
        ``<chain>:<resi>``) xor ``ref_ligand`` (``<chain>:<resi>`` or an SDF path), ``num_nodes_lig`` (tensor of sizes, drawn
        from the size prior if None), ``n_nodes_bias`` / ``n_nodes_min``, ``timesteps``, ``svdd`` (ATP) / ``spsa`` flags as
        in the reference; ``reward_fn`` replaces the reference's in-line RDKit scoring for the guided modes.
        Returns the list of molecules (``output.Molecule`` or whatever ``mol_builder`` returns; None results dropped)."""
        assert (pocket_ids is None) ^ (ref_ligand is None)
        pocket = self.pockets.get(pdb_file, pocket_ids, ref_ligand, repeats=n_samples)
        pocket_com_before = self._com(pocket['x'], pocket['mask'], n_samples)
        if num_nodes_lig is None:
            if self.size_distribution is None:
                raise ValueError('num_nodes_lig is None and no size histogram was given')
            num_nodes_lig = self.size_distribution.sample_conditional(n1=None, n2=pocket['size'])
        num_nodes_lig = torch.as_tensor(num_nodes_lig, device=self.device).long() + n_nodes_bias
        num_nodes_lig = torch.clamp(num_nodes_lig, min=n_nodes_min)
        xh_lig, xh_pocket, lig_mask, pocket_mask = self.ddpm.sample_given_pocket(
            pocket, num_nodes_lig, timesteps=timesteps, svdd=svdd, spsa=spsa, reward_fn=reward_fn, **kwargs)
        return self._finish(xh_lig, xh_pocket, lig_mask, pocket_mask, pocket_com_before, n_samples, sanitize,
                            largest_frag, relax_iter, return_tensors)

    def prepare_substructure(self, ref_ligand: str, fix_atoms: Sequence[str], pdb_model):
        """inpaint.py:47-62: the atoms to keep, as (coord [n,3] fp32, one_hot [n, atom_nf] int64).  ``fix_atoms`` is a list
        of SDF files (all atoms of each file's first molecule) or of atom names inside the PDB residue ``ref_ligand``."""
        enc = self.dataset_info['atom_encoder']
        if str(fix_atoms[0]).endswith('.sdf'):
            mols = [output.read_sdf(f)[0] for f in fix_atoms]
            coord = torch.cat([torch.from_numpy(m.positions).float() for m in mols], dim=0)
            types = torch.tensor([enc[a] for m in mols for a in m.symbols])
        else:
            chain, resi = ref_ligand.split(':')
            hits = [r for r in pdb_model if r.chain == chain and r.id[1] == int(resi)]
            assert len(hits) == 1
            atoms = [a for a in hits[0].get_atoms() if a.name in set(fix_atoms)]
            coord = torch.tensor([a.coord.tolist() for a in atoms], dtype=ingest.FLOAT_TYPE).reshape(-1, 3)
            types = torch.tensor([enc[a.element.capitalize()] for a in atoms], dtype=torch.long)
        return coord, torch.nn.functional.one_hot(types, num_classes=len(enc))

    @torch.no_grad()
    def inpaint_ligand(self, pdb_file, n_samples: int, ligand: str, fix_atoms: Sequence[str], add_n_nodes=None,
                       svdd: int = 0, center: str = 'ligand', sanitize: bool = False, largest_frag: bool = False,
                       relax_iter: int = 0, timesteps: Optional[int] = None, resamplings: int = 1,
                       reward_fn: Optional[Callable] = None, return_tensors: bool = False, **kwargs) -> List:
        """``inpaint.py:inpaint_ligand`` (inpaint.py:65-188) with the same arguments: ``ligand`` defines the pocket
        (``<chain>:<resi>`` or an SDF path), ``fix_atoms`` the substructure to keep, ``add_n_nodes`` how many atoms to add
        (drawn from the size prior, at least the fixed ones, if None).  ``save_traj`` (visualisation) is not offered."""
        dev = self.device
        pocket = self.pockets.get(pdb_file, None, ligand, repeats=n_samples)
        x_fixed, one_hot_fixed = self.prepare_substructure(ligand, fix_atoms, ingest.parse_pdb(pdb_file))
        n_fixed = len(x_fixed)
        if add_n_nodes is None:
            if self.size_distribution is None:
                raise ValueError('add_n_nodes is None and no size histogram was given')
            num_nodes_lig = torch.clamp(self.size_distribution.sample_conditional(n1=None, n2=pocket['size']), min=n_fixed)
        else:
            num_nodes_lig = torch.ones(n_samples, dtype=torch.long) * n_fixed + add_n_nodes
        num_nodes_lig = num_nodes_lig.to(dev)
        ligand_mask = ingest.num_nodes_to_batch_mask(len(num_nodes_lig), num_nodes_lig, dev)
        # fixed atoms occupy the first n_fixed slots of every sample (inpaint.py:125-139)
        starts = torch.cumsum(num_nodes_lig, 0) - num_nodes_lig
        slot = torch.arange(len(ligand_mask), device=dev) - starts[ligand_mask]
        is_fixed = slot < n_fixed
        lig = {'x': torch.zeros((len(ligand_mask), self.x_dims), device=dev, dtype=ingest.FLOAT_TYPE),
               'one_hot': torch.zeros((len(ligand_mask), self.atom_nf), device=dev, dtype=ingest.FLOAT_TYPE),
               'size': num_nodes_lig, 'mask': ligand_mask}
        lig['x'][is_fixed] = x_fixed.to(dev)[slot[is_fixed]]
        lig['one_hot'][is_fixed] = one_hot_fixed.to(dev, ingest.FLOAT_TYPE)[slot[is_fixed]]
        pocket_com_before = self._com(pocket['x'], pocket['mask'], n_samples)
        xh_lig, xh_pocket, lig_mask, pocket_mask = self.ddpm.inpaint(
            lig, pocket, is_fixed.long(), svdd=svdd, resamplings=resamplings, timesteps=timesteps, center=center,
            reward_fn=reward_fn, **kwargs)
        return self._finish(xh_lig, xh_pocket, lig_mask, pocket_mask, pocket_com_before, n_samples, sanitize,
                            largest_frag, relax_iter, return_tensors)

    def _finish(self, xh_lig, xh_pocket, lig_mask, pocket_mask, pocket_com_before, n_samples, sanitize, largest_frag,
                relax_iter, return_tensors):
        # move the generated molecules back to the original pocket position (:918-925)
        pocket_com_after = self._com(xh_pocket[:, :self.x_dims], pocket_mask, n_samples)
        shift = pocket_com_before - pocket_com_after
        xh_pocket[:, :self.x_dims] += shift[pocket_mask]
        xh_lig[:, :self.x_dims] += shift[lig_mask]
        x = xh_lig[:, :self.x_dims].contiguous()
        atom_type = xh_lig[:, self.x_dims:].argmax(1)
        if self.mol_builder is not None:            # host chemistry of the reference, per molecule (:935-947)
            xs, ts = x.cpu(), atom_type.cpu()
            mols = [self.mol_builder(px, pt, self.dataset_info, sanitize=sanitize, relax_iter=relax_iter,
                                     largest_frag=largest_frag)
                    for px, pt in zip(ingest.batch_to_list(xs, lig_mask.cpu()), ingest.batch_to_list(ts, lig_mask.cpu()))]
        else:
            mols = [output.process_molecule(m, add_hydrogens=False, sanitize=sanitize, relax_iter=relax_iter,
                                            largest_frag=largest_frag)
                    for m in output.build_molecules(x, atom_type, lig_mask, n_samples, self.dataset_info, self.perception)]
        mols = [m for m in mols if m is not None]
        if return_tensors:
            return mols, (xh_lig, xh_pocket, lig_mask, pocket_mask)
        return mols

    def generate_to_sdf(self, pdb_file, outfile, n_samples: int = 20, batch_size: Optional[int] = None,
                        num_nodes_lig: Optional[int] = None, all_frags: bool = False, **kw) -> int:
        """The body of the ``generate_ligands.py`` script (:92-108): ``n_samples // batch_size`` batches, largest fragment
        unless ``all_frags``, one SDF file.  Returns the number of molecules written."""
        batch_size = n_samples if batch_size is None else batch_size
        sizes = None if num_nodes_lig is None else torch.ones(batch_size, dtype=torch.long) * num_nodes_lig
        molecules = []
        for _ in range(n_samples // batch_size):
            molecules.extend(self.generate_ligands(pdb_file, batch_size, num_nodes_lig=sizes,
                                                   largest_frag=not all_frags, **kw))
        return output.write_sdf_file(outfile, molecules)
