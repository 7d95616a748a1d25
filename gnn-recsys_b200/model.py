"""``ConvModel`` and its layers -- the reference's ``src/model.py`` surface on the B200 kernels.

Same class names, constructor arguments, attribute names and ``state_dict`` keys as the reference
(``user_embed.proj_feats.{weight,bias}``, ``layers.{i}.mods.{etype}.{fc_self,fc_neigh,fc_preagg}.weight``), so
a checkpoint written by ``main_train.py:386`` loads unchanged and ``main_inference.py:102-121`` keeps working.
What differs is the execution: every ``ConvLayer`` is ONE fused kernel per relation (CSR gather-reduce +
``fc_self``/``fc_neigh`` + ReLU + L2 norm + cross-relation accumulate, ``gr_sage_relation_f32``) instead of DGL
``update_all`` + two sgemm + four elementwise kernels, and ``CosinePrediction`` is one gather-dot kernel per
etype (``gr_edge_cosine_f32``) instead of two ``F.normalize`` passes over every node row plus a DGL SDDMM.

Forward only: the kernels have no autograd (backward is outside the accelerated path).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N
from . import ops

_NN_AGGREGATORS = ('pool_nn', 'pool_nn_edge', 'mean_nn', 'mean_nn_edge')
_AGGREGATORS = ('mean', 'mean_nn', 'pool_nn', 'mean_edge', 'mean_nn_edge', 'pool_nn_edge')
_LSTM = ('lstm', 'lstm_edge')


class _TransposedWeight:
    """``weight.t().contiguous()`` cached until the parameter changes (kernels read weights k-major)."""

    def __init__(self):
        self._key, self._val = None, None

    def get(self, w: torch.Tensor) -> torch.Tensor:
        key = (w.data_ptr(), w._version, w.device, tuple(w.shape))
        if key != self._key:
            self._val = w.detach().to(torch.float32).t().contiguous()
            self._key = key
        return self._val


class NodeEmbedding(nn.Module):
    """Projects the node features into embedding space (reference ``src/model.py:10-24``)."""

    def __init__(self, in_feats, out_feats):
        super().__init__()
        self.proj_feats = nn.Linear(in_feats, out_feats)
        self._wt = _TransposedWeight()

    def forward(self, node_feats):
        w = self.proj_feats.weight
        x = node_feats.to(w.device, torch.float32)
        return ops.linear(x, self._wt.get(w), self.proj_feats.bias.detach())


class ConvLayer(nn.Module):
    """One layer of message passing and aggregation for one edge type (reference ``src/model.py:27-237``).

    ``forward(graph, x)`` takes the relation's CSR (a ``graph.Relation``) and ``x = (h_neigh, h_self)``.
    ``out`` / ``accumulate`` / ``z_scale`` let ``HeteroGraphConv`` fold its stack+reduce into the kernel's store.
    """

    def reset_parameters(self):
        gain = nn.init.calculate_gain('relu')
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)
        if self._aggre_type in _NN_AGGREGATORS:
            nn.init.xavier_uniform_(self.fc_preagg.weight, gain=gain)

    def __init__(self, in_feats: Tuple[int, int], out_feats: int, dropout: float, aggregator_type: str, norm):
        super().__init__()
        self._in_neigh_feats, self._in_self_feats = in_feats
        self._out_feats = out_feats
        self._aggre_type = aggregator_type
        self.dropout_fn = nn.Dropout(dropout)
        self.norm = norm
        self.fc_self = nn.Linear(self._in_self_feats, out_feats, bias=False)
        self.fc_neigh = nn.Linear(self._in_neigh_feats, out_feats, bias=False)
        if aggregator_type in _NN_AGGREGATORS:
            self.fc_preagg = nn.Linear(self._in_neigh_feats, self._in_neigh_feats, bias=False)
        if aggregator_type in _LSTM:
            raise NotImplementedError('lstm aggregators are outside the accelerated hot path (DESIGN.md, scope)')
        self._wt_self, self._wt_neigh, self._wt_pre = _TransposedWeight(), _TransposedWeight(), _TransposedWeight()
        self._packed_key, self._packed = None, None
        self.reset_parameters()

    def _packed_weights(self, wst: torch.Tensor, wnt: torch.Tensor):
        """Split / packed form of (fc_self, fc_neigh) for the fused kernel, rebuilt only when a weight changed."""
        ws, wn = self.fc_self.weight, self.fc_neigh.weight
        key = (ws.data_ptr(), ws._version, wn.data_ptr(), wn._version, str(ws.device))
        if key != self._packed_key:
            self._packed = ops.sage_pack_weights(wst, wnt)
            self._packed_key = key
        return self._packed

    def forward(self, graph, x, cetype=None, out=None, accumulate=N.ACC_STORE, z_scale=1.0, row_begin=0,
                row_end=None):
        h_neigh, h_self = x
        if self._aggre_type not in _AGGREGATORS:
            raise KeyError('Aggregator type {} not recognized.'.format(self._aggre_type))
        if self.training and self.dropout_fn.p > 0:
            h_neigh, h_self = self.dropout_fn(h_neigh), self.dropout_fn(h_self)
        base = self._aggre_type[:-5] if self._aggre_type.endswith('_edge') else self._aggre_type
        edge_w = None
        if self._aggre_type.endswith('_edge') and (
                cetype is None or (cetype[0] in ('user', 'item') and cetype[2] in ('user', 'item'))):
            if graph.weight is None:  # the reference reads graph.edata['occurrence'] (src/model.py:174)
                raise KeyError('occurrence')
            edge_w = graph.weight
        h_neigh = h_neigh.contiguous()
        if base in ('mean_nn', 'pool_nn'):  # messages = relu(fc_preagg(h)) for every source row (model.py:151,158)
            h_neigh = ops.linear(h_neigh, self._wt_pre.get(self.fc_preagg.weight), None, relu=True)
        reducer = N.REDUCE_MAX if base == 'pool_nn' else N.REDUCE_MEAN
        if out is None:
            out = torch.empty((h_self.shape[0], self._out_feats), dtype=torch.float32, device=h_self.device)
        wst, wnt = self._wt_self.get(self.fc_self.weight), self._wt_neigh.get(self.fc_neigh.weight)
        return ops.sage_relation(graph.indptr, graph.indices, edge_w, h_neigh, h_self.contiguous(), wst, wnt, out,
                                 reducer, bool(self.norm), accumulate, z_scale, row_begin, row_end,
                                 packed=self._packed_weights(wst, wnt))


class HeteroGraphConv(nn.Module):
    """``dgl.nn.pytorch.HeteroGraphConv`` (dgl 0.5.2) for blocks: one module per etype in ``self.mods``; relations
    without edges or without inputs are skipped; per destination type the results are reduced with
    ``aggregate`` in {'sum', 'mean', 'max'} -- folded into the kernels' accumulate mode, no stacking."""

    def __init__(self, mods, aggregate='sum'):
        super().__init__()
        self.mods = nn.ModuleDict(mods)
        if aggregate not in ('sum', 'mean', 'max'):
            raise KeyError(aggregate)
        self.aggregate = aggregate

    def forward(self, g, inputs, row_ranges=None, out_buffers=None):
        """``row_ranges`` (optional): ``{dst ntype: (begin, end)}`` -- compute only that destination shard
        (multi-GPU); rows outside it are left untouched in the returned (full-height) tensors. A block built by
        ``HeteroGraph.sharded_block_on`` (``g.shard_ranges``) holds only this rank's CSR rows: its ranges are used and
        the kernels walk the local ``indptr`` from 0. ``out_buffers`` (optional ``{dst ntype: [rows >= n, d_out]}``):
        write into these tensors (e.g. an all-gather buffer) instead of allocating."""
        shard = getattr(g, 'shard_ranges', None)
        if shard is not None:
            row_ranges = shard
        dst_inputs = {k: v[:g.number_of_dst_nodes(k)] for k, v in inputs.items()}
        todo: Dict[str, list] = {}
        for c in g.canonical_etypes:
            rel = g.rels[c]
            if rel.nnz == 0 or c[0] not in inputs or c[2] not in dst_inputs:
                continue
            todo.setdefault(c[2], []).append(c)
        rsts = {}
        for dtype, cs in todo.items():
            out = None if out_buffers is None else out_buffers.get(dtype)
            for i, c in enumerate(cs):
                last = i == len(cs) - 1
                acc = N.ACC_STORE if i == 0 else (N.ACC_MAX if self.aggregate == 'max' else N.ACC_ADD)
                scale = 1.0 / len(cs) if (self.aggregate == 'mean' and last) else 1.0
                rb, re = (0, None) if row_ranges is None else row_ranges[dtype]
                mod = self.mods[c[1]]
                if shard is None:
                    out = mod(g.rels[c], (inputs[c[0]], dst_inputs[dtype]), cetype=c, out=out, accumulate=acc,
                              z_scale=scale, row_begin=rb, row_end=re)
                else:  # local CSR rows [0, re - rb) <-> global destination rows [rb, re): views of the full tables
                    if out is None:
                        out = torch.empty((dst_inputs[dtype].shape[0], mod._out_feats), dtype=torch.float32,
                                          device=dst_inputs[dtype].device)
                    if re > rb:
                        mod(g.rels[c], (inputs[c[0]], dst_inputs[dtype][rb:re]), cetype=c, out=out[rb:re],
                            accumulate=acc, z_scale=scale)
            rsts[dtype] = out
        return rsts


class CosinePrediction(nn.Module):
    """Cosine similarity of the two end points of every edge to score (reference ``src/model.py:308-327``).
    Returns ``{canonical etype: [E, 1]}`` like ``graph.edata['cos']``; etypes whose node types carry no
    embedding are skipped (the reference's ``except KeyError: pass``)."""

    def forward(self, graph, h):
        ratings = {}
        for c in graph.canonical_etypes:
            if c[0] not in h or c[2] not in h:
                continue
            hs, hd = h[c[0]], h[c[2]]
            u, v = graph.device_edges(c, hs.device)
            ratings[c] = ops.edge_cosine(u, v, hs.contiguous(), hd.contiguous())
        return ratings


class ConvModel(nn.Module):
    """Embedding layers + ``ConvLayer`` stack + scoring function (reference ``src/model.py:330-470``)."""

    def __init__(self, g, n_layers: int, dim_dict, norm: bool = True, dropout: float = 0.0,
                 aggregator_type: str = 'mean', pred: str = 'cos', aggregator_hetero: str = 'sum',
                 embedding_layer: bool = True):
        super().__init__()
        self.embedding_layer = embedding_layer
        if embedding_layer:
            self.user_embed = NodeEmbedding(dim_dict['user'], dim_dict['hidden'])
            self.item_embed = NodeEmbedding(dim_dict['item'], dim_dict['hidden'])
            if 'sport' in g.ntypes:
                self.sport_embed = NodeEmbedding(dim_dict['sport'], dim_dict['hidden'])
        self.layers = nn.ModuleList()

        def hetero(in_of, out_dim):
            return HeteroGraphConv({etype[1]: ConvLayer(in_of(etype), out_dim, dropout, aggregator_type, norm)
                                    for etype in g.canonical_etypes}, aggregate=aggregator_hetero)

        if not embedding_layer:  # input layer on raw feature dims
            self.layers.append(hetero(lambda e: (dim_dict[e[0]], dim_dict[e[2]]), dim_dict['hidden']))
        for _ in range(n_layers - 2):  # hidden layers
            self.layers.append(hetero(lambda e: (dim_dict['hidden'], dim_dict['hidden']), dim_dict['hidden']))
        self.layers.append(hetero(lambda e: (dim_dict['hidden'], dim_dict['hidden']), dim_dict['out']))  # output layer
        if pred == 'cos':
            self.pred_fn = CosinePrediction()
        elif pred == 'nn':
            raise NotImplementedError("pred='nn' (MLP scorer) is outside the accelerated hot path (DESIGN.md, scope)")
        else:
            raise KeyError('Prediction function {} not recognized.'.format(pred))

    def get_repr(self, blocks, h, row_ranges=None):
        for i in range(len(blocks)):
            h = self.layers[i](blocks[i], h) if row_ranges is None else self.layers[i](blocks[i], h, row_ranges)
        return h

    def embed_type(self, ntype: str, feats):
        """NodeEmbedding of one node type's rows (any row subset: the multi-GPU path embeds each rank's own rows)."""
        return getattr(self, ntype + '_embed')(feats)

    def embed(self, h):
        """NodeEmbedding per node type, in place in the passed dict like the reference (``model.py:462-466``)."""
        h['user'] = self.user_embed(h['user'])
        h['item'] = self.item_embed(h['item'])
        if 'sport' in h.keys():
            h['sport'] = self.sport_embed(h['sport'])
        return h

    @torch.no_grad()
    def forward(self, blocks, h, pos_g, neg_g, embedding_layer: bool = True):
        dev = next(self.parameters()).device
        for k in list(h.keys()):
            h[k] = h[k].to(dev, torch.float32)
        if embedding_layer:
            self.embed(h)
        h = self.get_repr(blocks, h)
        pos_score = self.pred_fn(pos_g, h)
        neg_score = self.pred_fn(neg_g, h)
        return h, pos_score, neg_score


def max_margin_loss(pos_score, neg_score, delta: float, neg_sample_size: int, use_recency: bool = False,
                    recency_scores=None, remove_false_negative: bool = False, negative_mask=None, cuda=False,
                    device=None):
    """Max-margin loss over K-consecutive negatives per positive edge (reference ``src/model.py:473-533``).
    A handful of elementwise torch ops on the score vectors the kernels produced; left in PyTorch."""
    all_scores = None
    for etype in pos_score.keys():
        neg = neg_score[etype].reshape(-1, neg_sample_size)
        pos = pos_score[etype]
        if remove_false_negative:
            mask = negative_mask[etype].reshape(-1, neg_sample_size).to(neg.device)
        else:
            mask = torch.zeros(size=neg.shape, device=neg.device)
        scores = F.relu(neg + delta - pos - mask)
        if use_recency:
            try:
                scores = scores / torch.unsqueeze(recency_scores[etype].to(neg.device), 1)
            except KeyError:  # only training etypes carry recency
                pass
        all_scores = scores if all_scores is None else torch.cat((all_scores, scores), 0)
    if all_scores is None:
        all_scores = torch.empty(0)
    return torch.mean(all_scores)
