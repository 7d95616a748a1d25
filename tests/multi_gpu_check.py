"""Run under torchrun (one process per GPU): the id-range-sharded path must reproduce the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py
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
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    for agg, n_layers, hidden in (('mean', 2, 128), ('pool_nn', 3, 256)):
        data = grb.make_graph(3001, 1203, 60000, seed=5)
        g = data.graph()
        torch.manual_seed(7)
        model = grb.ConvModel(g, n_layers, {'user': 2, 'item': 4, 'hidden': hidden, 'out': 128}, True, 0.0, agg).to(dev).eval()
        blk = g.full_block_on(dev)
        blocks = [blk] * (n_layers - 1)
        feats = {t: g.nodes[t].data['features'].to(dev) for t in g.ntypes}
        with torch.no_grad():
            h1 = model.get_repr(blocks, model.embed(dict(feats)))
            hs = D.sharded_get_repr(model, blocks, model.embed(dict(feats)))
        for t in h1:
            assert torch.equal(h1[t], hs[t]), 'sharded embeddings differ for %s (%s)' % (t, agg)  # same kernels, same order
        with torch.no_grad():
            hb = D.sharded_get_repr(model, blocks, model.embed(dict(feats)), balance=('item', 'user'))
        for t in h1:
            assert torch.equal(h1[t], hb[t]), 'work-balanced sharding differs for %s (%s)' % (t, agg)
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
        assert bool(((a - b).abs() < 1e-5).all()), 'sharded top-k differs beyond score ties (%s)' % agg
        assert bool(((ids1[ub:ue] < 0) == (ids_s < 0)).all())
        ids_u, sc_u, (vb, ve) = D.sharded_recommend(hs['user'], hs['item'], 10, bought, item_shards=1)
        assert (vb, ve) == (ub, ue) and torch.equal(ids_u, ids1[ub:ue]), 'user-sharded layout differs (%s)' % agg
        ids_a, _, _ = D.sharded_recommend(hs['user'], hs['item'], 10, bought)  # default layout (shard the longer side)
        assert torch.equal(ids_a, ids_u)
        same = float((ids1[ub:ue] == ids_s).float().mean())
        if rank == 0:
            print('multi-gpu check ok: world=%d agg=%s identical ids %.4f (rest are ties < 1e-5)' % (world, agg, same))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
