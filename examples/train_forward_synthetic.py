#!/usr/bin/env python
"""The forward half of the reference's training loop (src/train/run.py:89-139 with the loaders of
src/sampling.py:153-207) on synthetic data and the B200 library: EdgeDataLoader batches -> ConvModel.forward ->
max_margin_loss. Forward only -- the kernels have no autograd (backward is outside the accelerated path).

    python examples/train_forward_synthetic.py --users 10000 --items 5000 --edges 200000 --neg 2500

`--host-loader` builds the blocks with the NumPy builder and moves them with block.to(device) like the reference does;
the default builds them on the GPU (device= on the loader). Both yield the same batches for the same seed.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnn_recsys_b200 as dgl  # noqa: E402  stands in for dgl + src.model
from gnn_recsys_b200 import ConvModel, max_margin_loss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--users', type=int, default=10000)
    ap.add_argument('--items', type=int, default=5000)
    ap.add_argument('--edges', type=int, default=200000)
    ap.add_argument('--batch', type=int, default=1024)      # fixed_params.edge_batch_size
    ap.add_argument('--neg', type=int, default=2500)        # params['neg_sample_size']
    ap.add_argument('--fanouts', default='10,10')
    ap.add_argument('--batches', type=int, default=20)
    ap.add_argument('--delta', type=float, default=0.266)
    ap.add_argument('--host-loader', action='store_true')
    args = ap.parse_args()
    device = torch.device('cuda:0')
    fanouts = [int(f) for f in args.fanouts.split(',')]

    data = dgl.make_graph(args.users, args.items, args.edges, seed=0)
    graph = data.graph()
    torch.manual_seed(1)
    model = ConvModel(graph, len(fanouts) + 1, {'user': 2, 'item': 4, 'hidden': 128, 'out': 128}, True, 0.0, 'mean',
                      'cos', 'sum', True).to(device).eval()
    train_eids = {'buys': np.arange(graph.num_edges('buys')), 'clicks': np.arange(graph.num_edges('clicks'))}
    loader = dgl.EdgeDataLoader(
        graph, train_eids, dgl.MultiLayerNeighborSampler(fanouts), exclude='reverse_types',
        reverse_etypes={'buys': 'bought-by', 'bought-by': 'buys', 'clicks': 'clicked-by', 'clicked-by': 'clicks'},
        negative_sampler=dgl.negative_sampler.Uniform(args.neg), batch_size=args.batch, shuffle=True, drop_last=False,
        seed=2, device=None if args.host_loader else device)

    total, t0 = 0.0, None
    for i, (_, pos_g, neg_g, blocks) in enumerate(loader):
        if i == 2:  # two warm-up batches
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        if i == args.batches + 2:
            break
        blocks = [b.to(device) for b in blocks]                       # run.py:104-107 (no-op for device-built blocks)
        input_features = blocks[0].srcdata['features']                # run.py:110
        _, pos_score, neg_score = model(blocks, input_features, pos_g, neg_g, True)   # run.py:118-122
        loss = max_margin_loss(pos_score, neg_score, args.delta, args.neg, cuda=True, device=device)   # run.py:123-133
        total += float(loss)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.batches
    print('%s loader: %.2f ms per training-step forward (%d positive + %d negative edges), mean loss %.5f'
          % ('host' if args.host_loader else 'device', dt * 1e3, args.batch, args.batch * args.neg,
             total / (args.batches + 2)))


if __name__ == '__main__':
    main()
