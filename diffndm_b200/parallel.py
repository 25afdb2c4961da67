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
