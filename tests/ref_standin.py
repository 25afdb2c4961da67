"""A structural stand-in for the reference's ``EGNNDynamics`` / ``ConditionalDDPM`` objects: the same attribute names and
``state_dict`` keys (dynamics.py:10-85, egnn_new.py:6-29, 69-95, 135-223, en_diffusion.py:40-60), no
forward.  ``/root/reference`` does not exist on the GPU box, so the drop-in constructors (``B200EGNNDynamics.from_reference``,
``B200ConditionalDDPM.from_reference``) are exercised there against this mirror; tests/test_from_reference.py checks on the
build container that the real reference module exposes exactly these attributes."""
import torch
from torch import nn


def _mlp(sizes, last_bias=True):
    layers = []
    for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:])):
        last = i == len(sizes) - 2
        layers.append(nn.Linear(a, b, bias=(last_bias or not last)))
        if not last:
            layers.append(nn.SiLU())
    return nn.Sequential(*layers)


class _GCL(nn.Module):
    def __init__(self, H, edges_in_d):
        super().__init__()
        self.attention = True
        self.edge_mlp = nn.Sequential(nn.Linear(2 * H + edges_in_d, H), nn.SiLU(), nn.Linear(H, H), nn.SiLU())
        self.node_mlp = nn.Sequential(nn.Linear(2 * H, H), nn.SiLU(), nn.Linear(H, H))
        self.att_mlp = nn.Sequential(nn.Linear(H, 1), nn.Sigmoid())


class _Equiv(nn.Module):
    def __init__(self, H, edges_in_d):
        super().__init__()
        self.tanh = True
        head = nn.Linear(H, 1, bias=False)                      # ONE object at the end of both MLPs, like the reference
        self.coord_mlp = nn.Sequential(nn.Linear(2 * H + edges_in_d, H), nn.SiLU(), nn.Linear(H, H), nn.SiLU(), head)
        self.cross_product_mlp = nn.Sequential(nn.Linear(2 * H + edges_in_d, H), nn.SiLU(), nn.Linear(H, H), nn.SiLU(), head)


class _Block(nn.Module):
    def __init__(self, H, edges_in_d, norm_constant, coords_range):
        super().__init__()
        self.n_layers = 1
        self.norm_constant = norm_constant
        self.coords_range_layer = coords_range
        self.add_module('gcl_0', _GCL(H, edges_in_d))
        self.add_module('gcl_equiv', _Equiv(H, edges_in_d))


class _EGNN(nn.Module):
    def __init__(self, in_nf, H, n_layers, norm_constant, normalization_factor, coords_range):
        super().__init__()
        self.hidden_nf, self.n_layers = H, n_layers
        self.normalization_factor = normalization_factor
        self.reflection_equiv = False
        self.embedding = nn.Linear(in_nf, H)
        self.embedding_out = nn.Linear(H, in_nf)
        for i in range(n_layers):
            self.add_module(f'e_block_{i}', _Block(H, 2, norm_constant, coords_range))


class StandInDynamics(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        A, R, J = cfg.atom_nf, cfg.residue_nf, cfg.joint_nf
        self.atom_encoder = _mlp([A, 2 * A, J])
        self.atom_decoder = _mlp([J, 2 * A, A])
        self.residue_encoder = _mlp([R, 2 * R, J])
        self.residue_decoder = _mlp([J, 2 * R, R])
        self.egnn = _EGNN(J + 1, cfg.hidden_nf, cfg.n_layers, cfg.norm_constant, cfg.normalization_factor, cfg.coords_range)
        self.n_dims = 3
        self.edge_cutoff_l, self.edge_cutoff_p, self.edge_cutoff_i = cfg.edge_cutoff_ligand, cfg.edge_cutoff_pocket, cfg.edge_cutoff_interaction
        self.edge_nf = 0 if cfg.edge_embedding_dim is None else cfg.edge_embedding_dim
        self.update_pocket_coords = False
        self.condition_time = True


class _Schedule(nn.Module):
    def __init__(self, table):
        super().__init__()
        self.gamma = nn.Parameter(table.clone(), requires_grad=False)


class StandInDDPM(nn.Module):
    def __init__(self, dynamics, gamma_table, timesteps=500):
        super().__init__()
        self.dynamics = dynamics
        self.gamma = _Schedule(gamma_table)
        self.T = timesteps
        self.norm_values = [1, 4]
        self.norm_biases = (None, 0)


def build(cfg, weights):
    m = StandInDynamics(cfg)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in weights.items()}, strict=True)
    return m.eval()
