"""Whole-call timing of the ``generate_ligands`` surface (SURVEY.md section 8f-3 / 8f-4 measurement): PDB file ->
pocket tensors -> 500-step sampling of B ligands -> bond perception -> molecules -> SDF, on one GPU.

Prints one JSON object with the wall-clock of every stage (median of three calls after a warm-up call):
  ingest_cold   parse + residue selection + upload (PocketCache miss)      ingest_hot   PocketCache hit
  sample        ConditionalSampler.sample_given_pocket (the hot path; CUDA-graph replay of the reverse step)
  molecules     dndm_bond_orders for the batch + host assembly + largest-fragment filter
  sdf           write_sdf_file
The synthetic pocket is written as a PDB file first (no network, no
reference data on the GPU box)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E, output, synthetic                  # noqa: E402
from diffndm_b200.datasets import crossdock_dataset_info                  # noqa: E402
from diffndm_b200.generate import LigandGenerator                         # noqa: E402
from diffndm_b200.sampler import ConditionalSampler                       # noqa: E402
from diffndm_b200.weights import DynamicsConfig, random_init              # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_p = int(sys.argv[2]) if len(sys.argv) > 2 else 330
T = int(sys.argv[3]) if len(sys.argv) > 3 else 500


def write_pdb(path, px, pt, lig_xyz):
    names = ['C', 'N', 'O', 'S']
    lines = []
    for i, (p, t) in enumerate(zip(px, pt)):
        el = names[min(int(t), 3)]
        lines.append(f"ATOM  {i + 1:5d}  {el + str(i % 8):<3s} ALA A{i // 8 + 1:4d}    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}"
                     f"  1.00  0.00          {el:>2s}")
    for k, p in enumerate(lig_xyz):
        lines.append(f"HETATM{len(px) + k + 1:5d}  C{k:<2d} LIG A 900    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}  1.00  0.00           C")
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\nEND\n')


def med(f, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), r


cfg = DynamicsConfig()
info = crossdock_dataset_info()
px, pt = synthetic.synthetic_pocket(7, n_p)
px = np.round(px, 3).astype(np.float32)
tmp = tempfile.mkdtemp()
pdb = os.path.join(tmp, 'pocket.pdb')
# ligand atoms spread through the pocket so that every residue is within 8 A (the whole synthetic pocket is selected)
write_pdb(pdb, px, pt, px[:: max(1, len(px) // 24)] * 0.6)
dyn = E.B200EGNNDynamics(cfg, random_init(cfg, 0, 1e-3), max_nodes=B * (n_p + 60) + 1024, max_edges=B * (n_p + 60) * 40,
                         max_samples=max(B, 8)).eval()
gen = LigandGenerator(ConditionalSampler(dyn, timesteps=500), info)
sizes = torch.from_numpy(synthetic.synthetic_ligand_sizes(7, B))

out = {'batch': B, 'pocket_atoms_requested': n_p, 'timesteps': T}


def cold():
    gen.pockets._entries.clear()
    return gen.pockets.get(pdb, ref_ligand='A:900', repeats=B)


cold()
out['ingest_cold_s'], pocket = med(cold)
out['ingest_hot_s'], pocket = med(lambda: gen.pockets.get(pdb, ref_ligand='A:900', repeats=B))
out['pocket_atoms'] = int(pocket['size'][0])

gen.ddpm.sample_given_pocket(pocket, sizes, timesteps=T)                      # warm-up (graph capture, allocator)
out['sample_s'], (xh_lig, xh_pocket, lig_mask, pocket_mask) = med(lambda: gen.ddpm.sample_given_pocket(pocket, sizes, timesteps=T))
# spread the atoms of the random-init trajectory back to chemical distances so that bond perception has work to do
x = xh_lig[:, :3].contiguous()
cnt = torch.bincount(lig_mask, minlength=B).float()
com = torch.zeros((B, 3), device=x.device).index_add_(0, lig_mask, x) / cnt[:, None]
rg = torch.zeros(B, device=x.device).index_add_(0, lig_mask, ((x - com[lig_mask]) ** 2).sum(1)).div(cnt).sqrt()
x = ((x - com[lig_mask]) * (2.5 / rg.clamp(min=1e-3))[lig_mask, None]).contiguous()
types = xh_lig[:, 3:].argmax(1)


def mols():
    ms = output.build_molecules(x, types, lig_mask, B, info, gen.perception)
    return [output.process_molecule(m, largest_frag=True) for m in ms]


mols()
out['molecules_s'], molecules = med(mols)
out['bonds_total'] = int(sum(m.GetNumBonds() for m in output.build_molecules(x, types, lig_mask, B, info, gen.perception)))
sdf = os.path.join(tmp, 'out.sdf')
out['sdf_s'], n = med(lambda: output.write_sdf_file(sdf, molecules))
out['sdf_bytes'] = os.path.getsize(sdf)

whole, _ = med(lambda: gen.generate_ligands(pdb, B, ref_ligand='A:900', num_nodes_lig=sizes, largest_frag=True, timesteps=T), n=2)
out['generate_ligands_s'] = whole
out['ligands_per_s_whole_call'] = B / whole * (500 / T)
out['ligands_per_s_sampling_only'] = B / out['sample_s'] * (500 / T)
out['flags'] = dyn.engine.read_flags()
print(json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in out.items()}))
