"""Seeded CrossDocked-shaped synthetic inputs and a BioPython-free pocket reader.

Synthetic pockets: N_p ~ clip(N(330, 80), 150, 700) heavy atoms, a jittered
minimum-spacing fill of a spherical shell 4-14 A around the origin so that the
pocket-pocket degree at 5 A is protein-like, coordinates rounded to 1e-3 A like
PDB files, atom types ~ Categorical(C .64, N .17, O .18, S .01) (constants.py:182
histogram collapsed to atoms), one-hot.  Ligands: n_i ~ clip(N(23, 8), 5, 50).
"""
from __future__ import annotations

import numpy as np

ATOM_TYPES = ['C', 'N', 'O', 'S', 'B', 'Br', 'Cl', 'P', 'I', 'F']   # constants.py:170-171
_POCKET_P = np.array([0.64, 0.17, 0.18, 0.01])


def synthetic_pocket(seed: int, n_atoms: int | None = None, min_spacing: float = 1.25,
                     r_in: float = 4.0, r_out: float | None = None, density: float = 0.04):
    """Return (x[N_p,3] float32 rounded to 1e-3, types[N_p] int64)."""
    rng = np.random.default_rng(1234 + seed)
    if n_atoms is None:
        n_atoms = int(np.clip(rng.normal(330, 80), 150, 700))
    # shell volume follows the atom count at pocket-like density (~0.04 atoms/A^3: pp degree ~17 at 5 A)
    if r_out is None:
        r_out = (3 * (n_atoms / density) / (4 * np.pi) + r_in ** 3) ** (1 / 3)
    pts = np.zeros((n_atoms, 3))
    k = 0
    tries = 0
    s2 = min_spacing ** 2
    while k < n_atoms and tries < 200:
        tries += 1
        cand = rng.normal(size=(4 * n_atoms, 3))
        cand /= np.linalg.norm(cand, axis=1, keepdims=True)
        r = (rng.uniform(r_in ** 3, r_out ** 3, size=(len(cand), 1))) ** (1 / 3)
        cand = cand * r
        for p in cand:
            if k == 0 or np.min(np.sum((pts[:k] - p) ** 2, axis=1)) >= s2:
                pts[k] = p
                k += 1
                if k >= n_atoms:
                    break
    x = np.round(pts[:n_atoms], 3).astype(np.float32)
    types = rng.choice(4, size=len(x), p=_POCKET_P).astype(np.int64)
    return x, types


def synthetic_ligand_sizes(seed: int, n_samples: int):
    rng = np.random.default_rng(4321 + seed)
    return np.clip(np.rint(rng.normal(23, 8, size=n_samples)), 5, 50).astype(np.int64)


_LIGAND_P = np.array([0.70, 0.12, 0.15, 0.01, 0.0, 0.0, 0.01, 0.0, 0.0, 0.01])    # C N O S B Br Cl P I F


def synthetic_ligand_pose(seed: int, lig_sizes, center, r_max: float | None = None, bond: float = 1.5,
                          min_spacing: float = 1.2, atom_nf: int = 10, norm_h: float = 4.0):
    """A ligand-shaped point cloud per sample in the pocket: a branched random walk with ``bond``-long steps that stays
    within ``r_max`` of ``center`` (default: 1.15 x the radius of a sphere holding n heavy atoms at the drug-like 18 A^3 per
    atom, at least 3.5 A -- real ligands are elongated; 3rfm's 14-atom ligand reaches 3.5 A from its centroid) and keeps
    ``min_spacing`` between atoms where it can.  Returns xh[N_l, 3+atom_nf] in the sampler's normalised units (absolute
    coordinates, one-hot / norm_h) -- the data point x_0 the synthetic score of bench.py pulls the reverse trajectory
    towards (random-init weights cannot denoise)."""
    rng = np.random.default_rng(777 + seed)
    out = []
    center = np.asarray(center, np.float64)
    r_fixed = r_max
    for n in np.asarray(lig_sizes):
        r_max = r_fixed if r_fixed is not None else max(3.5, 1.15 * (3.0 * 18.0 * float(n) / (4.0 * np.pi)) ** (1.0 / 3.0))
        pts = np.zeros((int(n), 3))
        pts[0] = rng.normal(size=3) * 0.8
        for i in range(1, int(n)):
            anchor = pts[rng.integers(max(0, i - 3), i)]          # grow from one of the last atoms: chains with short branches
            best, best_d = None, -1.0
            for _ in range(24):
                d = rng.normal(size=3)
                cand = anchor + bond * d / np.linalg.norm(d)
                if np.linalg.norm(cand) > r_max:
                    continue
                dm = np.min(np.linalg.norm(pts[:i] - cand, axis=1))
                if dm >= min_spacing:
                    best = cand
                    break
                if dm > best_d:
                    best, best_d = cand, dm
            if best is None:                                       # every try left the cavity: step towards the centre
                best = anchor * (1.0 - bond / max(np.linalg.norm(anchor), bond))
            pts[i] = best
        types = rng.choice(atom_nf, size=int(n), p=_LIGAND_P[:atom_nf] / _LIGAND_P[:atom_nf].sum())
        onehot = np.eye(atom_nf, dtype=np.float32)[types] / np.float32(norm_h)
        out.append(np.concatenate([(pts + center).astype(np.float32), onehot], axis=1))
    return np.concatenate(out, axis=0).astype(np.float32)


def make_batch(pocket_x, pocket_types, lig_sizes, seed: int, atom_nf: int = 10, norm_h: float = 4.0):
    """Assemble the reference's batch layout for one pocket repeated len(lig_sizes) times
    (prepare_pocket(repeats=n), lightning_modules.py:763-801; num_nodes_to_batch_mask,
    utils.py:145-153) and draw z_T ~ N(pocket COM, I) projected to the ligand-COM-free
    subspace (conditional_model.py:914-930).

    Returns dict with xh_lig[N_l,13], xh_pocket[N_p,13] (normalised: one-hot/4), lig_mask, pocket_mask."""
    rng = np.random.default_rng(99 + seed)
    B = len(lig_sizes)
    n_p = len(pocket_x)
    pocket_mask = np.repeat(np.arange(B, dtype=np.int64), n_p)
    lig_mask = np.repeat(np.arange(B, dtype=np.int64), lig_sizes)
    onehot_p = np.eye(atom_nf, dtype=np.float32)[pocket_types] / np.float32(norm_h)
    xh_pocket = np.concatenate([np.tile(pocket_x, (B, 1)), np.tile(onehot_p, (B, 1))], axis=1).astype(np.float32)
    com = pocket_x.mean(axis=0, dtype=np.float64).astype(np.float32)
    n_l = int(lig_sizes.sum())
    z = rng.normal(size=(n_l, 3 + atom_nf)).astype(np.float32)
    z[:, :3] += com
    # COM-free projection w.r.t. the ligand mean, applied to ligand and pocket
    sums = np.zeros((B, 3), np.float64)
    np.add.at(sums, lig_mask, z[:, :3].astype(np.float64))
    mean = (sums / lig_sizes[:, None]).astype(np.float32)
    z[:, :3] -= mean[lig_mask]
    xh_pocket[:, :3] -= mean[pocket_mask]
    return dict(xh_lig=z, xh_pocket=xh_pocket, lig_mask=lig_mask, pocket_mask=pocket_mask,
                lig_sizes=np.asarray(lig_sizes, np.int64), n_pocket=n_p)


# ----------------------------------------------------------------------------------------------
# minimal PDB / SDF readers (replace BioPython + RDKit for utils.get_pocket_from_ligand, utils.py:102-127)
# ----------------------------------------------------------------------------------------------
_STD_AA = {'ALA', 'ARG', 'ASN', 'ASP', 'CYS', 'GLN', 'GLU', 'GLY', 'HIS', 'ILE', 'LEU', 'LYS', 'MET',
           'PHE', 'PRO', 'SER', 'THR', 'TRP', 'TYR', 'VAL'}


def read_sdf_coords(path):
    with open(path) as f:
        lines = f.read().splitlines()
    n = int(lines[3][:3])
    return np.array([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in lines[4:4 + n]], np.float32)


def read_pocket_from_pdb(pdb_path, ligand_coords, dist_cutoff=8.0):
    """Residues (standard amino acids, first model) with any atom closer than ``dist_cutoff`` to the
    ligand; heavy atoms in the crossdock atom vocabulary (lightning_modules.py:773-783)."""
    residues = {}
    order = []
    with open(pdb_path) as f:
        for l in f:
            if l.startswith('ENDMDL'):
                break
            if not l.startswith('ATOM'):
                continue
            if l[16] not in (' ', 'A'):
                continue
            resname = l[17:20].strip()
            key = (l[21], int(l[22:26]), l[26])
            el = l[76:78].strip().capitalize() or l[12:16].strip()[0]
            xyz = (float(l[30:38]), float(l[38:46]), float(l[46:54]))
            if key not in residues:
                residues[key] = (resname, [])
                order.append(key)
            residues[key][1].append((el, xyz))
    xs, ts = [], []
    enc = {a: i for i, a in enumerate(ATOM_TYPES)}
    lig = np.asarray(ligand_coords, np.float32)
    for key in order:
        resname, atoms = residues[key]
        if resname not in _STD_AA:
            continue
        xyz = np.array([a[1] for a in atoms], np.float32)
        d = np.sqrt(((xyz[:, None] - lig[None]) ** 2).sum(-1))
        if d.min() < dist_cutoff:
            for el, p in atoms:
                if el == 'H':
                    continue
                if el in enc:
                    xs.append(p)
                    ts.append(enc[el])
    return np.array(xs, np.float32), np.array(ts, np.int64)


class PointMassScore:
    """``ConditionalSampler.eps_transform`` for benchmarks with random-init weights: adds the exact score of a point-mass data
    distribution at ``x0`` (a ligand pose per atom, relative to its sample's first pocket atom) to the network output,

        eps <- eps + (z_t - alpha_t (x0 + pocket_first_atom)) / sigma_t,

    so that a reverse trajectory converges to a ligand in the pocket like a trained model's does, instead of inflating by
    1 / alpha_T ~ 45x and leaving the pocket (no ligand-pocket edges: a degenerate, too cheap workload).  Device-side torch
    ops only (capture-safe); a batch that repeats the base batch (SPSA / ATP copies) reuses the rows modulo their count."""

    def __init__(self, x0_rel, gamma_table, n_pocket: int, timesteps: int, device):
        import torch
        self.torch = torch
        self.x0 = torch.as_tensor(x0_rel, dtype=torch.float32, device=device)
        g = torch.as_tensor(gamma_table, dtype=torch.float32, device=device)
        self.alpha = torch.sqrt(torch.sigmoid(-g))
        self.sigma = torch.sqrt(torch.sigmoid(g))
        self.n_p, self.T = int(n_pocket), int(timesteps)
        self._rows = {}

    def __call__(self, eps, z, xp, t, lig_mask, pocket_mask):
        torch = self.torch
        key = (z.shape[0], xp.shape[0])
        if key not in self._rows:              # index tensors per batch layout (built outside capture by the warm-up call)
            n = z.shape[0]
            self._rows[key] = (torch.arange(n, device=z.device) % self.x0.shape[0],
                               torch.arange(xp.shape[0] // self.n_p, device=z.device) * self.n_p)
            if len(self._rows) > 16:
                self._rows.pop(next(iter(self._rows)))
        rows, first = self._rows[key]
        idx = torch.round(t.reshape(-1) * self.T).long()
        a = self.alpha[idx][lig_mask][:, None]
        s = self.sigma[idx][lig_mask][:, None]
        tgt = self.x0[rows].clone()
        tgt[:, :3] += xp[first][lig_mask, :3]
        return eps + (z - a * tgt) / s
