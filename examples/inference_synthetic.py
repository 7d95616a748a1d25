#!/usr/bin/env python
"""main_inference.py of the reference (main_inference.py:20-175), on synthetic data and the B200 library.

Same call sequence: graph -> ConvModel (+ state_dict) -> full-neighbour NodeDataLoader -> get_embeddings ->
create_already_bought -> get_recs -> precision / recall / coverage. Run on a machine with a B200:

    python examples/inference_synthetic.py --users 100000 --items 20000 --edges 3000000
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnn_recsys_b200 as dgl  # noqa: E402  stands in for dgl + src.model + src.train.run + src.metrics
from gnn_recsys_b200 import ConvModel, get_embeddings, get_recs, create_already_bought, get_metrics_at_k  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--users', type=int, default=20000)
    ap.add_argument('--items', type=int, default=5000)
    ap.add_argument('--edges', type=int, default=400000)
    ap.add_argument('--k', type=int, default=10)
    ap.add_argument('--n-layers', type=int, default=3)
    ap.add_argument('--aggregator', default='mean_nn')
    args = ap.parse_args()
    device = torch.device('cuda:0')

    data = dgl.make_graph(args.users, args.items, args.edges, seed=0)
    graph = dgl.heterograph(data.relations(), {'user': args.users, 'item': args.items})   # src/builder.py:382
    graph.nodes['user'].data['features'] = data.user_feat                                 # src/utils_data.py:241-317
    graph.nodes['item'].data['features'] = data.item_feat

    dim_dict = {'user': 2, 'item': 4, 'hidden': 128, 'out': 128}
    torch.manual_seed(1)
    model = ConvModel(graph, args.n_layers, dim_dict, True, 0.0, args.aggregator, 'cos', 'sum', True).to(device)
    model.eval()   # a trained checkpoint loads with model.load_state_dict(torch.load(path)) -- same keys as the reference

    user_ids = np.arange(args.users)
    sampler = dgl.MultiLayerFullNeighborSampler(args.n_layers - 1)
    loader = dgl.NodeDataLoader(graph, {'user': user_ids, 'item': np.arange(args.items)}, sampler, batch_size=128, shuffle=True,
                                drop_last=False, num_workers=0)   # main_inference.py:130-138
    t0 = time.perf_counter()
    embeddings = get_embeddings(graph, dim_dict['out'], model, loader, len(loader), True, device, True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bought_eids = graph.out_edges(u=torch.from_numpy(user_ids), form='eid', etype='buys')
    already_bought = create_already_bought(graph, bought_eids)
    sample = user_ids[:1000].tolist()
    recs = get_recs(graph, embeddings, model, dim_dict['out'], args.k, sample, already_bought, True, True, device)
    t2 = time.perf_counter()
    print('embeddings %s / %s in %.3f s; %d users recommended in %.3f s' % (tuple(embeddings['user'].shape),
                                                                            tuple(embeddings['item'].shape), t1 - t0,
                                                                            len(recs), t2 - t1))
    print('user 0 ->', [int(i) for i in recs[0]])
    # metrics against a made-up "future purchases" ground truth (every user's clicked items)
    clicks = data.relations()[('user', 'clicks', 'item')]
    p, r, c = get_metrics_at_k(embeddings, graph, model, dim_dict['out'], (clicks[0], clicks[1]), bought_eids, args.k, True,
                               True, device)
    print('precision@%d %.4f  recall@%d %.4f  coverage %.4f' % (args.k, p, args.k, r, c))


if __name__ == '__main__':
    main()
