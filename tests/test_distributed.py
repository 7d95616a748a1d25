"""world_size-2 gloo runs of the multi-GPU host logic on the CPU: id-range sharding, the per-layer all-gather of
embedding rows and the routing + merge of per-shard top-k lists (the kernels themselves are covered by -m gpu)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import straightline as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_users, n_items, k, ret):
    import gnn_recsys_b200 as grb
    D = grb.distributed
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        # ---- all-gather of a table of which each rank computed only its own id range
        full = torch.randn(n_users, 8, generator=g)
        b, e = D.shard_range(n_users, world, rank)
        local = torch.full((n_users, 8), float('nan'))
        local[b:e] = full[b:e]
        out = D.allgather_rows(local, n_users)
        assert torch.equal(out, full)
        # unequal contiguous ranges (work-balanced sharding)
        bounds = [0, n_users // 5, n_users] if world == 2 else [0] + [n_users * (r + 1) // world for r in range(world)]
        local = torch.full((n_users, 8), float('nan'))
        local[bounds[rank]:bounds[rank + 1]] = full[bounds[rank]:bounds[rank + 1]]
        assert torch.equal(D.allgather_rows_v(local, bounds), full)
        # ---- per-shard exact top-k of ALL users (oracle on this rank's item range) -> owner-side merge
        hu = torch.nn.functional.normalize(torch.rand(n_users, 16, generator=g), dim=1)
        hi = torch.nn.functional.normalize(torch.rand(n_items, 16, generator=g), dim=1)
        ib, ie = D.shard_range(n_items, world, rank)
        scores = hu @ hi[ib:ie].t()
        kk = min(k, ie - ib)
        val, idx = torch.topk(scores, kk, dim=1)
        ids = torch.full((n_users, k), -1, dtype=torch.int32)
        sc = torch.full((n_users, k), float('-inf'))
        ids[:, :kk], sc[:, :kk] = (idx + ib).to(torch.int32), val
        all_ids, all_sc, ub, ue = D.exchange_topk(ids, sc)
        assert all_ids.shape == (world, D.chunk_rows(n_users, world), k)
        ms, mi = O.merge_partial_topk([all_sc[p].numpy() for p in range(world)], [all_ids[p].numpy() for p in range(world)], k)
        want = torch.topk(hu[ub:ue] @ hi.t(), k, dim=1).indices.numpy()
        assert np.array_equal(mi[:ue - ub], want), (rank, mi[:3], want[:3])
        ret[rank] = (ub, ue)
    finally:
        dist.destroy_process_group()


def _run(world, n_users, n_items, k):
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_users, n_items, k, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return dict(ret)


def test_two_ranks_allgather_and_topk_merge():
    ret = _run(2, 101, 57, 10)  # odd sizes: padded chunks
    assert ret[0] == (0, 51) and ret[1] == (51, 101)


def test_shard_ranges_cover_and_are_contiguous():
    import gnn_recsys_b200 as grb
    D = grb.distributed
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            rs = [D.shard_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert all(e - b <= D.chunk_rows(n, world) for b, e in rs)


def test_balanced_bounds_equalise_work():
    import gnn_recsys_b200 as grb
    D = grb.distributed
    d = grb.make_graph(500, 200, 20000, seed=2)
    d.items[:4000] = 3  # a hub item with 20 % of all edges
    blk = d.graph().full_block()
    for world in (2, 4, 8):
        b = D.balanced_bounds(blk, 'item', world)
        assert b[0] == 0 and b[-1] == 200 and all(x <= y for x, y in zip(b, b[1:]))
        deg = sum((blk.rels[c].indptr[1:] - blk.rels[c].indptr[:-1]).long() for c in blk.rels if c[2] == 'item') + 8
        work = [int(deg[b[r]:b[r + 1]].sum()) for r in range(world)]
        equal = [int(deg[s:e].sum()) for s, e in (D.shard_range(200, world, r) for r in range(world))]
        assert max(work) <= max(equal)
        assert max(work) <= max(int(deg.max()), int(deg.sum()) // world + int(deg.max()))
