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
    ok = ok and _packed_atp_case(rank, world)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def _packed_atp_case(rank, world):
    """atp_select_packed: the candidate groups of ONE pocket split over the ranks (group g on rank g % world), one
    all-gather per event, every rank rebuilds the winners of the reference's global top-B over index g * B + i."""
    from diffndm_b200.parallel import atp_select_packed
    B, G, n_p, D = 3, 5, 4, 13
    g = torch.Generator().manual_seed(1)
    sizes = torch.tensor([2, 5, 3])
    lig_mask = torch.repeat_interleave(torch.arange(B), sizes)
    pocket_mask = torch.repeat_interleave(torch.arange(B), n_p)
    n_l = int(sizes.sum())
    xh_pocket = torch.randn(B * n_p, D, generator=g)
    z_all = torch.randn(G, n_l, D, generator=g)                       # latents of every group
    shift_all = torch.randn(G, B, 3, generator=g)                     # every candidate's pocket translation
    scores = torch.randn(G * B, generator=g)
    scores[7] = scores[2]                                             # a tie: the lower candidate index wins
    p_all = xh_pocket.reshape(1, B, n_p, D).repeat(G, 1, 1, 1)
    p_all[..., :3] += shift_all[:, :, None, :]
    mine = [q for q in range(G) if q % world == rank]
    z, xp, m = atp_select_packed(torch.cat([scores[q * B:(q + 1) * B] for q in mine]), torch.cat([z_all[q] for q in mine]),
                                 torch.cat([p_all[q].reshape(B * n_p, D) for q in mine]), lig_mask, pocket_mask, xh_pocket,
                                 B, G, None)
    order = torch.sort(scores, descending=True, stable=True).indices[:B]
    ez, ep, em = [], [], []
    for pos, c in enumerate(order.tolist()):
        q, i = c // B, c % B
        ez.append(z_all[q][lig_mask == i])
        ep.append(p_all[q, i])
        em += [pos] * int(sizes[i])
    return torch.equal(z, torch.cat(ez)) and torch.allclose(xp, torch.cat(ep), atol=1e-6) and m.tolist() == em


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
