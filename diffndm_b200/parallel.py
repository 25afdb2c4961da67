"""Multi-GPU plumbing: one process per GPU, pockets sharded across ranks, no collective on the denoising path.

The only exchange the algorithm has is the ATP ("SVDD") selection when ONE pocket's candidate groups are split
across ranks (reference conditional_model.py:1203-1232 takes a global top-k over all candidates): every rank
contributes the scores and latent states of its candidates, and all ranks rebuild the same winners.  The payload
is ~100 candidates x ~23 atoms x 13 floats (~120 KB), so a plain ``all_gather`` (NCCL over NVLink on GPUs, gloo in
the CPU tests) is latency-bound and sufficient.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_pockets(n_pockets: int, rank: int, world: int) -> List[int]:
    """Static round-robin of pocket ids over ranks (work unit = one pocket's trajectory batch)."""
    return list(range(rank, n_pockets, world))


class PocketQueue:
    """Shared work queue over pocket ids for tail balance: pocket cost grows with its edge count
    (~ atoms x degree), so a static round-robin leaves ranks idle at the end of a job.  Every rank pulls the next pocket
    with an atomic add on the process group's key-value store -- host-side, a few bytes per pocket, no collective and
    nothing on the denoising path.  With ``costs`` the queue hands out the most expensive pockets first (longest
    processing time first), the order every rank derives identically from the same list.

    ``store=None`` uses the default process group's store; without an initialised process group the queue is local."""

    def __init__(self, n_pockets: int, costs: Optional[Sequence[float]] = None, store=None, key: str = 'dndm/pocket_queue'):
        self.n = int(n_pockets)
        if costs is not None:
            assert len(costs) == self.n
            self.order = sorted(range(self.n), key=lambda i: (-float(costs[i]), i))
        else:
            self.order = list(range(self.n))
        if store is None and dist.is_available() and dist.is_initialized():
            store = dist.distributed_c10d._get_default_store()
        self.store = store
        self.key = key
        self._local = 0

    def _next(self) -> int:
        if self.store is None:
            self._local += 1
            return self._local - 1
        return int(self.store.add(self.key, 1)) - 1

    def __iter__(self) -> Iterator[int]:
        while True:
            i = self._next()
            if i >= self.n:
                return
            yield self.order[i]


def _all_gather_varlen(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    """all_gather of tensors whose first dimension differs per rank (pad to the max, trim after)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    nmax = int(max(int(x) for x in ns))
    pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[: int(k)] for o, k in zip(out, ns)]


def atp_select_distributed(scores: torch.Tensor, z_lig: torch.Tensor, lig_sizes: torch.Tensor, top_k: int,
                           group=None, per_candidate: torch.Tensor = None):
    """Global top-k over candidates that live on different ranks.

    scores    [C_r]        mixed reward of this rank's candidates
    z_lig     [sum n_i,13] their latent ligand states, candidate-major
    lig_sizes [C_r]        atoms per candidate
    Returns (z_sel [sum n_sel, 13], mask_sel [sum n_sel] with values 0..top_k-1, sizes_sel [top_k]) -- identical on
    every rank, winners ordered by decreasing score with ties broken by global candidate index (deterministic).
    ``per_candidate`` [C_r, k] (optional) is any fixed-width per-candidate payload (e.g. the candidate's pocket
    translation and source sample); the winners' rows are returned as a fourth value.
    """
    all_scores = torch.cat(_all_gather_varlen(scores.float(), group))
    all_sizes = torch.cat(_all_gather_varlen(lig_sizes.long(), group))
    all_z = torch.cat(_all_gather_varlen(z_lig.float(), group))
    # stable descending sort == topk with a deterministic tie rule
    order = torch.sort(all_scores, descending=True, stable=True).indices[:top_k]
    starts = torch.cumsum(all_sizes, 0) - all_sizes
    zs, ms = [], []
    for rank_pos, c in enumerate(order.tolist()):
        s, n = int(starts[c]), int(all_sizes[c])
        zs.append(all_z[s:s + n])
        ms.append(torch.full((n,), rank_pos, dtype=torch.long, device=all_z.device))
    if per_candidate is not None:
        all_pc = torch.cat(_all_gather_varlen(per_candidate.float(), group))
        return torch.cat(zs), torch.cat(ms), all_sizes[order], all_pc[order]
    return torch.cat(zs), torch.cat(ms), all_sizes[order]


def atp_select_packed(mixed: torch.Tensor, big_z: torch.Tensor, big_p: torch.Tensor, lig_mask: torch.Tensor,
                      pocket_mask: torch.Tensor, xh_pocket: torch.Tensor, B: int, G: int, group=None):
    """ATP selection when the candidate groups of one pocket are split over the ranks of ``group`` (group g lives on rank
    g % world; the ranks carry the same pre-event state).  ONE all-gather per event: every rank contributes a fixed-size
    block [slots, B + 3 B + N_l * D] -- the mixed scores of its candidates, the ABSOLUTE position of each candidate's
    first pocket atom (the pocket is rigid: that is its whole state) and the candidate latents -- padded with -inf scores
    for the slots it does not own.  All ranks then rebuild the same winners, ordered like the reference's global top-k over
    candidate index g * B + i (conditional_model.py:1203-1232; ties: lower index first).

    mixed [n_here * B], big_z [n_here * N_l, D], big_p [n_here * N_p, D]: this rank's candidates, its groups in increasing
    order.  Returns (z_lig, xh_pocket, lig_mask) of the winners."""
    dev = xh_pocket.device
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_l, D = lig_mask.shape[0], xh_pocket.shape[1]
    n_p = xh_pocket.shape[0] // B
    slots = (G + world - 1) // world                                       # groups per rank, at most
    mine = [g for g in range(G) if g % world == rank]
    blk = B + 3 * B + n_l * D
    buf = torch.zeros((slots, blk), device=dev)
    buf[:, :B] = float('-inf')
    first = torch.arange(B, device=dev) * n_p
    for j, g in enumerate(mine):
        buf[j, :B] = mixed[j * B:(j + 1) * B]
        buf[j, B:4 * B] = big_p[j * B * n_p:(j + 1) * B * n_p][first, :3].reshape(-1)
        buf[j, 4 * B:] = big_z[j * n_l:(j + 1) * n_l].reshape(-1)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)                                 # the one collective of the event
    out = torch.stack(outs)
    # group g sits in block (g % world, g // world)
    gi = torch.arange(G, device=dev)
    blocks = out[gi % world, gi // world]                                   # [G, blk], candidate-index order
    scores = blocks[:, :B].reshape(-1)
    order = torch.sort(scores, descending=True, stable=True).indices[:B]
    g_sel, i_sel = order // B, order % B
    sizes = torch.bincount(lig_mask, minlength=B)
    starts = torch.cumsum(sizes, 0) - sizes
    sel_sizes = sizes[i_sel]
    total = int(sel_sizes.sum())
    new_m = torch.repeat_interleave(torch.arange(B, device=dev), sel_sizes, output_size=total)
    new_starts = torch.cumsum(sel_sizes, 0) - sel_sizes
    row = starts[i_sel][new_m] + (torch.arange(total, device=dev) - new_starts[new_m])
    lat = blocks[:, 4 * B:].reshape(G, n_l, D)
    z_new = lat[g_sel[new_m], row].contiguous()
    # winners' pockets: the local pocket of the source sample as a rigid template, moved to the winner's absolute position
    pos = blocks[:, B:4 * B].reshape(G, B, 3)[g_sel, i_sel]                  # [B,3] first-atom positions
    tmpl = xh_pocket.reshape(B, n_p, D)[i_sel].clone()                     # [B, n_p, D]
    tmpl[:, :, :3] += (pos - tmpl[:, 0, :3])[:, None, :]
    return z_new, tmpl.reshape(B * n_p, D).contiguous(), new_m
