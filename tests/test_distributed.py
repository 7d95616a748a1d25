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


def _cpu_kernels(grb):
    """Stand-ins for the three C-ABI calls the embedding pass makes, in plain torch on the CPU (oracle semantics), so
    that the SHARDING logic around them (local CSR rows, row ranges, all-gather buffers) can run under gloo."""
    ops = grb.ops

    def csr_build(src, dst, n_dst, validate=True):
        a = O.csr_by_dst(src.numpy().astype(np.int64), dst.numpy().astype(np.int64), n_dst)
        return tuple(torch.from_numpy(x) for x in a)

    def linear(x, wt, bias=None, relu=False):
        y = x.float() @ wt + (bias if bias is not None else 0)
        return torch.relu(y) if relu else y

    def sage_relation(indptr, indices, edge_w, h_src, h_dst, w_self_t, w_neigh_t, out, reducer, l2norm, accumulate=0,
                      z_scale=1.0, row_begin=0, row_end=None, flags=0, packed=None):
        n = indptr.shape[0] - 1
        row_end = n if row_end is None else row_end
        deg = (indptr[1:] - indptr[:-1]).long()
        dst = torch.repeat_interleave(torch.arange(n), deg)
        msg = h_src[indices.long()] * (edge_w[:, None] if edge_w is not None else 1.0)
        agg = torch.zeros(n, h_src.shape[1])
        if reducer == grb._native.REDUCE_MAX:
            agg = torch.full((n, h_src.shape[1]), float('-inf')).scatter_reduce(0, dst[:, None].expand_as(msg), msg, 'amax')
            agg[deg == 0] = 0
        else:
            agg.index_add_(0, dst, msg)
            agg = agg / deg.clamp(min=1)[:, None]
        z = torch.relu(h_dst[:n] @ w_self_t + agg @ w_neigh_t)
        if l2norm:
            nrm = z.norm(dim=1, keepdim=True)
            z = z / torch.where(nrm == 0, torch.ones_like(nrm), nrm)
        r = slice(row_begin, row_end)
        if accumulate == grb._native.ACC_ADD:
            z = out[:n] + z
        elif accumulate == grb._native.ACC_MAX:
            z = torch.maximum(out[:n], z)
        out[r] = (z * z_scale)[r]
        return out

    ops.csr_build, ops.linear, ops.sage_relation = csr_build, linear, sage_relation
    ops.sage_pack_weights = lambda *a, **k: None


def _worker_sharded(rank, world, port, ret):
    """sharded storage (HeteroGraph.sharded_block_on + distributed.sharded_forward) == the un-sharded pass, 3-layer
    pool_nn and 2-layer mean with hetero 'mean', odd node counts (padded chunks, a short last shard)."""
    import gnn_recsys_b200 as grb
    D = grb.distributed
    _cpu_kernels(grb)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        d = grb.make_graph(203, 77, 3000, seed=5)
        g = d.graph()
        num = {'user': 203, 'item': 77}
        for n_layers, agg, hetero in ((3, 'pool_nn', 'sum'), (2, 'mean', 'mean')):
            torch.manual_seed(7)
            model = grb.ConvModel(g, n_layers, {'user': 2, 'item': 4, 'hidden': 16, 'out': 8}, True, 0.0, agg, 'cos', hetero).eval()
            with torch.no_grad():
                feats = {t: g.nodes[t].data['features'] for t in g.ntypes}
                want = model.get_repr([g.full_block()] * (n_layers - 1), model.embed(dict(feats)))
                ranges = D.node_ranges(num, world, rank)
                sblk = g.sharded_block_on('cpu', ranges)
                assert sblk.shard_ranges == ranges
                for c, r in sblk.rels.items():
                    b, e = ranges[c[2]]
                    assert r.n_dst == e - b and r.indptr.shape[0] == e - b + 1 and int(r.indptr[0]) == 0
                    full_ip = g.csr(c)[0]
                    assert int(r.indptr[-1]) == int(full_ip[e] - full_ip[b])          # only this rank's edges are stored
                    assert np.array_equal(r.eperm.numpy(), g.csr(c)[2][full_ip[b]:full_ip[e]])  # same stable slot order
                local = {t: feats[t][ranges[t][0]:ranges[t][1]] for t in feats}
                got = D.sharded_forward(model, [sblk] * (n_layers - 1), local)
                for t in want:
                    assert torch.allclose(got[t], want[t], rtol=1e-6, atol=1e-7), (t, agg)
                # items cut by WORK (unequal ranges), users by rows
                wb = {'item': D.work_bounds(g, 'item', world)}
                assert wb['item'][0] == 0 and wb['item'][-1] == 77
                r2 = D.node_ranges(num, world, rank, wb)
                sb2 = g.sharded_block_on('cpu', r2, bounds=wb)
                loc2 = {t: feats[t][r2[t][0]:r2[t][1]] for t in feats}
                got2 = D.sharded_forward(model, [sb2] * (n_layers - 1), loc2)
                for t in want:
                    assert torch.allclose(got2[t], want[t], rtol=1e-6, atol=1e-7), (t, agg, 'work bounds')
                part = D.sharded_forward(model, [sblk] * (n_layers - 1), local, gather_last=('item',))
                ub, ue = ranges['user']
                assert torch.allclose(part['user'][ub:ue], want['user'][ub:ue], rtol=1e-6, atol=1e-7)
                assert torch.allclose(part['item'], want['item'], rtol=1e-6, atol=1e-7)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_ranks_sharded_storage_forward_equals_unsharded():
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


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
