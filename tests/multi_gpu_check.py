"""Run under torchrun (one process per rank): the id-range-sharded path must reproduce the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

One GPU per rank over NCCL when the box has enough GPUs; otherwise (GR_CHECK_BACKEND=gloo) the ranks share the GPUs and
the collectives go through gloo -- the same kernels, shards and exchange logic, so the driver's single-GPU test box
exercises the whole multi-rank path too.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnn_recsys_b200 as grb  # noqa: E402
D = grb.distributed


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    backend = os.environ.get('GR_CHECK_BACKEND', 'nccl')
    dev = torch.device('cuda', local % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    if backend == 'nccl':
        dist.init_process_group('nccl', device_id=dev)
    else:
        dist.init_process_group('gloo')
    world, rank = dist.get_world_size(), dist.get_rank()
    for agg, n_layers, hidden in (('mean', 2, 128), ('pool_nn', 3, 256)):
        data = grb.make_graph(3001, 1203, 60000, seed=5)
        data.items[:5000] = 11  # a hub item (> 2048 in-edges: the multi-CTA long-row path inside a shard)
        g = data.graph()
        num = {'user': data.n_users, 'item': data.n_items}
        torch.manual_seed(7)
        model = grb.ConvModel(g, n_layers, {'user': 2, 'item': 4, 'hidden': hidden, 'out': 128}, True, 0.0, agg).to(dev).eval()
        blk = g.full_block_on(dev)
        blocks = [blk] * (n_layers - 1)
        feats = {t: g.nodes[t].data['features'].to(dev) for t in g.ntypes}
        with torch.no_grad():
            h1 = model.get_repr(blocks, model.embed(dict(feats)))
            # (1) sharded STORAGE: this rank's CSR rows and feature rows only
            ranges = D.node_ranges(num, world, rank)
            sblk = g.sharded_block_on(dev, ranges)
            assert sum(r.nnz for r in sblk.rels.values()) < sum(r.nnz for r in blk.rels.values()) or world == 1
            local_feats = {t: feats[t][ranges[t][0]:ranges[t][1]].contiguous() for t in feats}
            hs = D.sharded_forward(model, [sblk] * (n_layers - 1), local_feats)
            hp = D.sharded_forward(model, [sblk] * (n_layers - 1), local_feats, gather_last=('item',))
            # (2) sharded COMPUTE over a replicated graph (row_begin / row_end of the kernels), equal and balanced rows
            hr = D.sharded_get_repr(model, blocks, model.embed(dict(feats)))
            hb = D.sharded_get_repr(model, blocks, model.embed(dict(feats)), balance=('item', 'user'))
        for t in h1:  # same kernels, same summation order: bit-identical
            assert torch.equal(h1[t], hs[t]), 'sharded-storage embeddings differ for %s (%s)' % (t, agg)
            assert torch.equal(h1[t], hr[t]), 'sharded embeddings differ for %s (%s)' % (t, agg)
            assert torch.equal(h1[t], hb[t]), 'work-balanced sharding differs for %s (%s)' % (t, agg)
        ub, ue = ranges['user']
        assert torch.equal(hp['user'][ub:ue], h1['user'][ub:ue]) and torch.equal(hp['item'], h1['item'])
        bi = D.balanced_bounds(blk, 'item', world)
        assert bi[0] == 0 and bi[-1] == data.n_items and all(a <= b for a, b in zip(bi, bi[1:]))
        buys = data.relations()[('user', 'buys', 'item')]
        bought = grb.BoughtCSR.from_edges(buys[0], buys[1], data.n_users)
        ids1, sc1 = grb.recommend_topk(h1['user'], grb.ScoringTable(h1['item'], grb.RecsConfig()), 10, bought)
        ids_s, sc_s, (ub, ue) = D.sharded_recommend(hs['user'], hs['item'], 10, bought, item_shards=world)
        assert ids_s.shape[0] == ue - ub
        hu = torch.nn.functional.normalize(h1['user'][ub:ue], dim=1)
        hi = torch.nn.functional.normalize(h1['item'], dim=1)
        a = (hu.unsqueeze(1) * hi[ids1[ub:ue].long().clamp(min=0)]).sum(-1)
        b = (hu.unsqueeze(1) * hi[ids_s.long().clamp(min=0)]).sum(-1)
        assert bool(((a - b).abs() < 1e-5).all()), 'item-sharded top-k differs beyond score ties (%s)' % agg
        assert bool(((ids1[ub:ue] < 0) == (ids_s < 0)).all())
        ids_u, sc_u, (vb, ve) = D.sharded_recommend(hp['user'], hp['item'], 10, bought, item_shards=1)
        c = (hu.unsqueeze(1) * hi[ids_u.long().clamp(min=0)]).sum(-1)
        assert (vb, ve) == (ub, ue) and bool(((a - c).abs() < 1e-5).all()), 'user-sharded layout differs (%s)' % agg
        ids_a, _, _ = D.sharded_recommend(hs['user'], hs['item'], 10, bought)  # default layout (shard the longer side)
        assert torch.equal(ids_a, ids_u)
        same = float((ids1[ub:ue] == ids_s).float().mean())
        if rank == 0:
            print('multi-gpu check ok: world=%d backend=%s agg=%s identical ids %.4f (rest are ties < 1e-5)'
                  % (world, backend, agg, same))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
