"""``get_embeddings`` -- the reference's ``src/train/run.py:311-349`` on the B200 kernels.

Same signature and return value (``{ntype: FloatTensor[num_nodes(ntype), out_dim]}``, zero rows for nodes that were
not seeded). The reference walks ``ceil((U + I) / 128)`` sampled mini-batches, re-expanding shared neighbourhoods
in every one; with a full-neighbour sampler the embedding of a node does not depend on the batch it is computed
in, so a loader in full-graph mode (``NodeDataLoader(..., batch_size=None)``) is served by ONE layer-wise pass over
the device-resident CSR of the whole graph. Mini-batch loaders still work (same loop as the reference).
"""
from __future__ import annotations

import torch


def _is_arange(ids: torch.Tensor, n: int) -> bool:
    return ids.numel() == n and (n == 0 or (int(ids[0]) == 0 and int(ids[-1]) == n - 1 and
                                            bool((ids[1:] - ids[:-1] == 1).all())))


@torch.no_grad()
def get_embeddings(g, out_dim: int, trained_model, nodeloader_test, num_batches_valid: int, cuda: bool = False,
                   device=None, embedding_layer: bool = True):
    """Fetch the embeddings for all the nodes in the nodeloader (see module docstring).

    ``cuda`` / ``device`` keep the reference's meaning for where the RESULT lives; the computation itself always
    runs on a CUDA device (``device`` or the model's device) -- there is no CPU path.
    """
    dev = torch.device(device) if device is not None else next(trained_model.parameters()).device
    if dev.type != 'cuda':
        dev = torch.device('cuda', torch.cuda.current_device())  # raises without a GPU: no CPU fallback
    trained_model = trained_model.to(dev)
    y = {}
    i2 = 0
    full = getattr(nodeloader_test, 'full_graph', False)
    for input_nodes, output_nodes, blocks in nodeloader_test:
        i2 += 1
        if i2 % 10 == 0:
            print("Computing embeddings: Batch {} out of {}".format(i2, num_batches_valid))
        if full:
            blk = g.full_block_on(dev, getattr(nodeloader_test, 'edge_weight', None))
            blocks = [blk] * len(blocks)
            input_features = {t: g.nodes[t].data['features'].to(dev, torch.float32, non_blocking=True)
                              for t in g.ntypes if 'features' in g.nodes[t].data}
            output_nodes = {t: v for t, v in output_nodes.items()}
        else:
            blocks = [b.to(dev) for b in blocks]
            input_features = {t: v.to(dev, torch.float32) for t, v in blocks[0].srcdata['features'].items()}
        if embedding_layer:
            input_features['user'] = trained_model.user_embed(input_features['user'])
            input_features['item'] = trained_model.item_embed(input_features['item'])
            if 'sport' in input_features.keys():
                input_features['sport'] = trained_model.sport_embed(input_features['sport'])
        h = trained_model.get_repr(blocks, input_features)
        for ntype in h.keys():
            if ntype not in output_nodes:  # full-graph pass: a type nobody seeded (e.g. 'sport') keeps its zero rows
                continue
            ids = output_nodes[ntype]
            n = g.num_nodes(ntype)
            if full and _is_arange(ids, n):
                y[ntype] = h[ntype]  # every row seeded: the layer output IS the table
                continue
            if ntype not in y:
                y[ntype] = torch.zeros(n, out_dim, device=dev)
            idx = ids.to(dev)
            y[ntype][idx] = h[ntype][idx] if full else h[ntype]
    for ntype in g.ntypes:  # node types that were never reached keep zero rows (run.py:329-333)
        if ntype not in y:
            y[ntype] = torch.zeros(g.num_nodes(ntype), out_dim, device=dev)
    if not cuda:
        y = {t: v.cpu() for t, v in y.items()}
    return y
