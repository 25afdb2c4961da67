"""Host-side scoring pool for the guided samplers.

In the reference every SPSA / ATP event scores its candidate molecules one after the other on the sampling process
(``handle_to_mol`` -> ``build_molecule`` -> ``MoleculeProperties``; conditional_model.py:845-882, analysis/metrics.py:282-368),
and with the denoiser two orders of magnitude faster that serial loop is what bounds guided sampling.  ``PooledReward``
keeps the scoring itself untouched -- any picklable ``score_one(positions [n,3] float32, atom_types [n] int64) -> float``
(RDKit / OpenBabel code in production) -- and changes only where it runs: the candidates leave the GPU in ONE copy,
are split per molecule and scored by worker processes in parallel; results come back in candidate order.  It plugs in
wherever the samplers take ``reward_fn(x_lig, atom_types, lig_mask) -> list[float]``.
"""
from __future__ import annotations

import multiprocessing as mp
from concurrent.futures import ProcessPoolExecutor
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch


def _score_chunk(args):
    score_one, mols = args
    return [float(score_one(x, t)) for x, t in mols]


class PooledReward:
    def __init__(self, score_one: Callable[[np.ndarray, np.ndarray], float], workers: int = 8, chunk: Optional[int] = None,
                 start_method: str = 'spawn'):
        """``chunk``: molecules per worker task; None = ceil(n / (2 * workers)) per call, i.e. two equal rounds over the
        workers (a fixed chunk size leaves workers idle in the last round: 17 tasks on 8 workers take 3 rounds)."""
        self.score_one = score_one
        self.workers = int(workers)
        self.chunk = None if chunk is None else int(chunk)
        self._pool: Optional[ProcessPoolExecutor] = None
        self._copy_stream = None
        if self.workers > 0:
            self._pool = ProcessPoolExecutor(max_workers=self.workers, mp_context=mp.get_context(start_method))

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True, cancel_futures=True)
            self._pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def split(x: np.ndarray, types: np.ndarray, mask: np.ndarray):
        """Per-molecule views of a batch whose mask is sorted (utils.py:145-153 layout)."""
        n = int(mask.max()) + 1 if len(mask) else 0
        bounds = np.searchsorted(mask, np.arange(n + 1))
        return [(x[bounds[i]:bounds[i + 1]], types[bounds[i]:bounds[i + 1]]) for i in range(n)]

    def _to_host(self, x_lig: torch.Tensor, atom_types: torch.Tensor, lig_mask: torch.Tensor, after=None):
        """One device->host transfer for the whole candidate set (coordinates, types and mask packed side by side).
        ``after``: a CUDA event; the copy then runs on this pool's side stream as soon as that event has fired, without
        waiting for work queued behind it on the producing stream."""
        if x_lig.is_cuda and after is not None:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=x_lig.device)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(after)
                packed = torch.cat([x_lig[:, :3].float(), atom_types.reshape(-1, 1).float(), lig_mask.reshape(-1, 1).float()],
                                   dim=1)
                host = torch.empty(packed.shape, dtype=packed.dtype, pin_memory=True)
                host.copy_(packed, non_blocking=True)
                for t in (x_lig, atom_types, lig_mask, packed):
                    t.record_stream(self._copy_stream)
            self._copy_stream.synchronize()
            host = host.numpy()
        else:
            packed = torch.cat([x_lig[:, :3].float(), atom_types.reshape(-1, 1).float(), lig_mask.reshape(-1, 1).float()], dim=1)
            host = packed.detach().to('cpu', non_blocking=False).numpy()
        x = np.ascontiguousarray(host[:, :3], dtype=np.float32)
        types = host[:, 3].astype(np.int64)
        mask = host[:, 4].astype(np.int64)
        return self.split(x, types, mask)

    def submit(self, x_lig: torch.Tensor, atom_types: torch.Tensor, lig_mask: torch.Tensor, after=None) -> 'PendingScores':
        """Non-blocking variant of ``__call__``: copies the candidates to the host (after the CUDA event ``after`` if given)
        and hands them to the workers; ``.result()`` of the returned handle gives the scores in candidate order.  The
        samplers use it to score one half of an SPSA round while the GPU denoises the other half."""
        mols = self._to_host(x_lig, atom_types, lig_mask, after)
        if self._pool is None:
            return PendingScores(None, _score_chunk((self.score_one, mols)))
        step = self.chunk or max(1, -(-len(mols) // (2 * self.workers)))
        chunks = [mols[i:i + step] for i in range(0, len(mols), step)]
        return PendingScores([self._pool.submit(_score_chunk, (self.score_one, c)) for c in chunks], None)

    def __call__(self, x_lig: torch.Tensor, atom_types: torch.Tensor, lig_mask: torch.Tensor) -> List[float]:
        return self.submit(x_lig, atom_types, lig_mask).result()


class PendingScores:
    def __init__(self, futures, ready):
        self._futures, self._ready = futures, ready

    def result(self) -> List[float]:
        if self._ready is None:
            out: List[float] = []
            for f in self._futures:
                out.extend(f.result())
            self._ready = out
        return self._ready


def radius_of_gyration_score(x: np.ndarray, types: np.ndarray) -> float:
    """A chemistry-free stand-in used by the tests and the guided-mode benchmark: compact molecules score higher."""
    if len(x) == 0:
        return 0.0
    return -float(np.sqrt(((x - x.mean(0)) ** 2).sum(1).mean()))
