"""CPU, world_size 2 over gloo: pocket sharding and the distributed ATP selection (the only exchange on the path)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from diffndm_b200.parallel import atp_select_distributed, shard_pockets
    g = torch.Generator().manual_seed(0)
    # 5 candidates in total, ragged: rank 0 owns 3, rank 1 owns 2
    sizes_all = torch.tensor([4, 7, 5, 6, 3])
    scores_all = torch.tensor([0.3, 0.9, 0.9, 0.1, 0.5])
    z_all = torch.randn(int(sizes_all.sum()), 13, generator=g)
    own = [0, 1, 2] if rank == 0 else [3, 4]
    starts = torch.cumsum(sizes_all, 0) - sizes_all
    z_own = torch.cat([z_all[int(starts[c]):int(starts[c] + sizes_all[c])] for c in own])
    z, m, s = atp_select_distributed(scores_all[own], z_own, sizes_all[own], top_k=3)
    # expected winners: candidates 1, 2 (tie broken by index), then 4
    exp = torch.cat([z_all[int(starts[c]):int(starts[c] + sizes_all[c])] for c in (1, 2, 4)])
    ok = torch.equal(z, exp) and s.tolist() == [7, 5, 3] and m.tolist() == [0] * 7 + [1] * 5 + [2] * 3
    # fixed-width per-candidate payload (pocket translation + source sample) follows the winners; an empty rank is legal
    pay_all = torch.arange(5 * 4, dtype=torch.float32).reshape(5, 4)
    z2, m2, s2, pay = atp_select_distributed(scores_all[own], z_own, sizes_all[own], top_k=2, per_candidate=pay_all[own])
    ok = ok and torch.equal(pay, pay_all[[1, 2]]) and s2.tolist() == [7, 5]
    if rank == 0:
        z3, m3, s3 = atp_select_distributed(scores_all, z_all, sizes_all, top_k=2)
    else:
        z3, m3, s3 = atp_select_distributed(scores_all[:0], z_all[:0], sizes_all[:0], top_k=2)
    ok = ok and s3.tolist() == [7, 5] and torch.equal(z3, exp[:12])
    ok = ok and shard_pockets(7, rank, world) == list(range(rank, 7, world))
    # shared work queue: every pocket handed out exactly once over both ranks, most expensive first
    from diffndm_b200.parallel import PocketQueue
    costs = [3.0, 9.0, 1.0, 9.0, 5.0, 2.0, 7.0]
    q = PocketQueue(7, costs=costs)
    mine = []
    for pid in q:
        mine.append(pid)
        if rank == 1:
            import time
            time.sleep(0.02)                     # a slow rank ends up with fewer pockets
    got = [None, None]
    dist.all_gather_object(got, mine)
    ok = ok and sorted(got[0] + got[1]) == list(range(7)) and q.order == [1, 3, 6, 4, 0, 5, 2]
    ok = ok and all(q.order.index(a) < q.order.index(b) for g in got for a, b in zip(g, g[1:]))
    ok = ok and len(got[0]) >= len(got[1])
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_pocket_queue_local():
    from diffndm_b200.parallel import PocketQueue
    assert list(PocketQueue(4)) == [0, 1, 2, 3]
    assert list(PocketQueue(3, costs=[1, 5, 5])) == [1, 2, 0]
    assert list(PocketQueue(0)) == []


def test_atp_selection_world2_gloo():
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
