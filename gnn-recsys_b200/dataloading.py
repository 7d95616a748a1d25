"""Block construction on the host: samplers and loaders feeding the hot path.

Mirrors the ``dgl.dataloading`` surface the reference uses (``main_inference.py:126-138``,
``src/sampling.py:153-241``): ``MultiLayerFullNeighborSampler``, ``MultiLayerNeighborSampler``,
``NodeDataLoader``, ``EdgeDataLoader`` and ``negative_sampler.Uniform``, so that the reference's call sites keep
working and config 4 (fan-out [10, 10] blocks + 1024 positive / 1024*K negative edges) has inputs.

Two builders behind the same classes:
  * the host builder in this file (NumPy) -- the reference's call shape: CPU loader, then ``block.to(device)``;
  * ``device=`` on a loader: frontiers, negatives and ``to_block`` compaction run as CUDA kernels on the graph
    structure resident in HBM (``sampling_device.py``; ``gr_sample_*``, ``gr_negative_uniform_i64``,
    ``gr_remap_first_appearance_i64``) and the batch never touches the host.
Both draw ONE 63-bit key per batch from the loader's generator; everything else is counter based
(``hash64(key, edge id)``, see include/gnn_recsys_b200.h), so the two builders produce the same blocks bit for bit.

Block layout follows DGL's ``to_block``: destination nodes = seeds in the given order; source nodes =
the destination nodes first, then unseen edge sources in first-appearance order (canonical etype
order, then destination row, then edge id -- the order of the block's CSR). ``NodeDataLoader(...,
batch_size=None)`` -- the fast path -- yields a single full-graph block per layer instead of ``ceil(n/128)``
sampled mini-batches; for a full-neighbour sampler the embeddings of the seed nodes are identical (SURVEY.md 8a,
row a12).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from .graph import Block, HeteroGraph, Relation, NID, EID, _as_np_ids


def _first_appearance_unique(arrs: List[np.ndarray]) -> np.ndarray:
    cat = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.int64)
    if cat.size == 0:
        return cat.astype(np.int64)
    uniq, first = np.unique(cat, return_index=True)
    return uniq[np.argsort(first, kind='stable')].astype(np.int64)


def _relabel(ids: np.ndarray, space: np.ndarray) -> np.ndarray:
    """Position of each ``ids`` value inside ``space`` (``space`` holds unique values, any order)."""
    if ids.size == 0:
        return ids.astype(np.int64)
    order = np.argsort(space, kind='stable')
    pos = np.searchsorted(space[order], ids)
    return order[pos].astype(np.int64)


_M64 = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15


def _fin64(x: int) -> int:
    x ^= x >> 30
    x = (x * 0xbf58476d1ce4e5b9) & _M64
    x ^= x >> 27
    x = (x * 0x94d049bb133111eb) & _M64
    return x ^ (x >> 31)


def sample_key(seed: int, stream_id: int) -> int:
    """``gr_sample_key``: the 64-bit key of one random stream (layer x relation, or negatives of one etype)."""
    return _fin64((_fin64((int(seed) + _GOLDEN) & _M64) + (int(stream_id) + 1) * _GOLDEN) & _M64)


def hash64(key: int, ctr: np.ndarray) -> np.ndarray:
    """Vectorised ``hash64(key, ctr)`` of include/gnn_recsys_b200.h (uint64 wrap-around arithmetic)."""
    with np.errstate(over='ignore'):
        x = (np.asarray(ctr).astype(np.uint64) + np.uint64(1)) * np.uint64(_GOLDEN) + np.uint64(key)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return x


NEGATIVE_STREAM = 4096  # stream ids: layer * 64 + relation index for frontiers, NEGATIVE_STREAM + relation index


def _draw_key(rng) -> int:
    return int(rng.integers(0, 2 ** 63, dtype=np.int64))


AUTO = 'auto'
_TORCH_LOADER_KWARGS = ('num_workers', 'pin_memory', 'collate_fn', 'persistent_workers', 'prefetch_factor', 'timeout',
                        'worker_init_fn', 'generator', 'use_ddp', 'ddp_seed')  # accepted and ignored: no worker processes


def resolve_edge_weight(g: HeteroGraph, edge_weight):
    """DGL blocks carry the parent graph's edge frames, so ``graph.edata['occurrence']`` (read by the ``*_edge``
    aggregators, ``src/model.py:174``) is there whenever the graph has it. ``'auto'`` (the loaders' default) mirrors
    that: attach ``'occurrence'`` when any relation carries it. ``None`` attaches nothing, a name attaches that field."""
    if edge_weight != AUTO:
        return edge_weight
    return 'occurrence' if any('occurrence' in g.edges[c].data for c in g.canonical_etypes) else None


def _check_loader_kwargs(kind: str, kwargs):
    unknown = [k for k in kwargs if k not in _TORCH_LOADER_KWARGS]
    if unknown:
        raise TypeError('%s got unexpected keyword argument(s) %s' % (kind, ', '.join(sorted(unknown))))


class LazyFullBlock:
    """Placeholder for the full-graph block of a loader in full-graph mode. ``get_embeddings`` swaps it for the
    device-resident block (``HeteroGraph.full_block_on``) without ever building the HOST CSR (a stable argsort over
    every edge list plus int64 copies: minutes and GBs at 500M edges, for a block nobody reads). Any other consumer
    gets the real thing: ``.to(device)`` returns the device block, any other attribute materialises the host block."""

    def __init__(self, g: HeteroGraph, edge_weight):
        self._g, self._edge_weight, self._host = g, edge_weight, None

    def to(self, device, *args, **kwargs):
        if torch.device(device).type == 'cuda':
            return self._g.full_block_on(torch.device(device), self._edge_weight)
        return self.materialize()

    def materialize(self) -> Block:
        if self._host is None:
            self._host = self._g.full_block(self._edge_weight)
        return self._host

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class _FrontierSampler:
    """Shared block builder; subclasses choose which in-edges of the seeds form the frontier."""

    def __init__(self, num_layers: int):
        self.num_layers = num_layers

    def frontier(self, g: HeteroGraph, seeds: Dict[str, np.ndarray], layer: int, key: int, exclude=None):
        """Per canonical etype: (src ids, dst ids, edge ids) of the chosen in-edges, edge-id order."""
        out = {}
        for ci, c in enumerate(g.canonical_etypes):
            s, d = g.edge_arrays(c)
            sd = seeds.get(c[2])
            if sd is None or sd.size == 0 or s.size == 0:
                out[c] = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64))
                continue
            indptr, _, eperm = g.csr(c)
            rows = np.unique(sd)
            lo, hi = indptr[rows].astype(np.int64), indptr[rows + 1].astype(np.int64)
            deg = hi - lo
            slots = np.repeat(lo - np.cumsum(deg) + deg, deg) + np.arange(int(deg.sum()))
            eids = eperm[slots].astype(np.int64)
            if exclude is not None and c in exclude and exclude[c].size:
                keep = ~np.isin(eids, exclude[c])
                eids = eids[keep]
                row_of = np.repeat(np.arange(rows.size), deg)[keep]
                deg = np.bincount(row_of, minlength=rows.size)
            fan = self._fanout(layer)
            if fan is not None and eids.size:
                row_of = np.repeat(np.arange(rows.size), deg)
                ticket = hash64(sample_key(key, layer * 64 + ci), eids) >> np.uint64(32)
                order = np.lexsort((eids, ticket, row_of))  # in a row edge-id order == CSR-slot order
                start = np.cumsum(deg) - deg
                rank = np.arange(eids.size) - np.repeat(start, deg)
                eids = eids[order][rank < fan]
            eids = np.sort(eids)
            out[c] = (s[eids].astype(np.int64), d[eids].astype(np.int64), eids)
        return out

    def _fanout(self, layer: int):
        return None

    def sample_blocks(self, g: HeteroGraph, seed_nodes, rng=None, exclude=None, edge_weight: Optional[str] = AUTO,
                      key: Optional[int] = None, device=None) -> List[Block]:
        """Blocks for ``seed_nodes``, innermost layer first. ``key`` (or one draw from ``rng``) fixes every random
        choice; ``device`` = a CUDA device builds the blocks there (``seed_nodes`` / ``exclude`` may then be device
        tensors) with the same result. ``edge_weight``: see ``resolve_edge_weight``."""
        edge_weight = resolve_edge_weight(g, edge_weight)
        if key is None:
            key = _draw_key(rng if rng is not None else np.random.default_rng(0))
        if device is not None and torch.device(device).type == 'cuda':
            from .sampling_device import sample_blocks_device
            return sample_blocks_device(g, self, seed_nodes, key, torch.device(device), exclude, edge_weight)
        seeds = {t: _as_np_ids(v).astype(np.int64) for t, v in seed_nodes.items()}
        if exclude is not None:
            exclude = {c: _as_np_ids(v).astype(np.int64) for c, v in exclude.items()}
        blocks: List[Block] = []
        for layer in reversed(range(self.num_layers)):
            fr = self.frontier(g, seeds, layer, key, exclude)
            blocks.insert(0, to_block(g, fr, seeds, edge_weight))
            seeds = {t: blocks[0].srcnodes[t].data[NID].numpy() for t in blocks[0].srctypes}
        return blocks


def to_block(g: HeteroGraph, frontier, dst_nodes: Dict[str, np.ndarray], edge_weight: Optional[str] = None) -> Block:
    """Compact a frontier into a ``Block`` (DGL ``to_block`` semantics, see module docstring)."""
    csr = {}
    for c, (s, d, eids) in frontier.items():  # CSR order of every relation: destination row, then edge id
        dn = dst_nodes.get(c[2], np.zeros(0, np.int64))
        ld = _relabel(d, dn)
        order = np.argsort(ld, kind='stable')
        csr[c] = (s[order], ld, order)
    src_ids = {}
    for t in g.ntypes:
        parts = [dst_nodes[t]] if t in dst_nodes else []
        parts += [csr[c][0] for c in frontier if c[0] == t]
        src_ids[t] = _first_appearance_unique(parts)
    rels = {}
    for c, (s, d, eids) in frontier.items():
        dn = dst_nodes.get(c[2], np.zeros(0, np.int64))
        ld, order = csr[c][1], csr[c][2]
        ls = _relabel(s, src_ids[c[0]])
        indptr = np.zeros(dn.size + 1, dtype=np.int64)
        np.cumsum(np.bincount(ld, minlength=dn.size), out=indptr[1:])
        w = None
        if edge_weight is not None and edge_weight in g.edges[c].data:
            w = g.edges[c].data[edge_weight].detach().cpu().to(torch.float32).reshape(-1)[
                torch.from_numpy(eids[order])].contiguous()
        rels[c] = Relation(torch.from_numpy(indptr.astype(np.int32)), torch.from_numpy(ls[order].astype(np.int32)),
                           int(src_ids[c[0]].size), int(dn.size), torch.from_numpy(eids[order]), w)
    num_src = {t: int(src_ids[t].size) for t in g.ntypes}
    num_dst = {t: int(dst_nodes[t].size) if t in dst_nodes else 0 for t in g.ntypes}
    sf, df = {}, {}
    for t in g.ntypes:
        sid = torch.from_numpy(src_ids[t])
        did = torch.from_numpy(dst_nodes[t]) if t in dst_nodes else torch.zeros(0, dtype=torch.int64)
        sf[t] = {k: v[sid] for k, v in g.nodes[t].data.items()}
        sf[t][NID] = sid
        df[t] = {k: v[did] for k, v in g.nodes[t].data.items()}
        df[t][NID] = did
    return Block(rels, num_src, num_dst, sf, df)


class MultiLayerFullNeighborSampler(_FrontierSampler):
    """``dgl.dataloading.MultiLayerFullNeighborSampler`` (``main_inference.py:129``)."""


class MultiLayerNeighborSampler(_FrontierSampler):
    """``dgl.dataloading.MultiLayerNeighborSampler(fanouts, replace=False)`` (``src/sampling.py:159``):
    at most ``fanouts[layer]`` in-edges per seed node and relation, drawn uniformly without replacement."""

    def __init__(self, fanouts, replace=False):
        super().__init__(len(fanouts))
        if replace:
            raise NotImplementedError('replace=True is not used by the reference')
        self.fanouts = list(fanouts)

    def _fanout(self, layer):
        return self.fanouts[layer]


class NodeDataLoader:
    """``dgl.dataloading.NodeDataLoader`` surface: iterate ``(input_nodes, output_nodes, blocks)``.

    With a full-neighbour sampler the embedding of a seed node does not depend on the batch it is computed in, so
    the loader takes the fast path whatever ``batch_size`` says (the reference passes 128, ``main_inference.py:134``):
    ONE iteration whose blocks are (lazy) full-graph blocks; ``get_embeddings`` then keeps only the seeded rows.
    ``force_minibatch=True`` restores the reference's batching (sampled blocks per ``batch_size`` seeds).
    KNOWN DIVERGENCE of the fast path (README, "Drop-in notes"): DGL's ``HeteroGraphConv`` skips a relation that has
    no edge in the CURRENT mini-batch, which removes that relation's ``fc_self`` term for every node of the batch; in
    one full-graph pass the relation is only skipped when the whole graph has no such edge. The two agree unless a
    128-node batch happens to contain no edge of some relation -- an artefact of batching that
    the golden case ``tiny_mean_batched`` reproduces with ``force_minibatch=True``.
    """

    def __init__(self, g: HeteroGraph, nids, block_sampler, batch_size=None, shuffle=False, drop_last=False,
                 seed=0, edge_weight=AUTO, force_minibatch=False, device=None, **kwargs):
        _check_loader_kwargs('NodeDataLoader', kwargs)
        self.g, self.sampler = g, block_sampler
        self.device = device
        self.nids = {t: _as_np_ids(v).astype(np.int64) for t, v in nids.items()}
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.rng = np.random.default_rng(seed)
        self.edge_weight = resolve_edge_weight(g, edge_weight)
        self.force_minibatch = force_minibatch
        self._flat_t = np.concatenate([np.full(v.size, i) for i, (t, v) in enumerate(sorted(self.nids.items()))]) \
            if self.nids else np.zeros(0, np.int64)
        self._flat_i = np.concatenate([v for _, v in sorted(self.nids.items())]) if self.nids else np.zeros(0, np.int64)
        self._types = [t for t, _ in sorted(self.nids.items())]

    @property
    def full_graph(self) -> bool:
        return isinstance(self.sampler, MultiLayerFullNeighborSampler) and (
            not self.force_minibatch or self.batch_size is None or self.batch_size >= self._flat_i.size)

    def __len__(self):
        n = self._flat_i.size
        if self.batch_size is None or self.full_graph:
            return 1
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        if self.full_graph:
            blk = LazyFullBlock(self.g, self.edge_weight)
            blocks = [blk] * self.sampler.num_layers
            ids = {t: torch.arange(self.g.num_nodes(t)) for t in self.g.ntypes}
            yield ids, {t: torch.from_numpy(v) for t, v in self.nids.items()}, blocks
            return
        n = self._flat_i.size
        order = self.rng.permutation(n) if self.shuffle else np.arange(n)
        for b in range(len(self)):
            sel = order[b * self.batch_size:(b + 1) * self.batch_size]
            seeds = {}
            for ti, t in enumerate(self._types):
                ids = self._flat_i[sel[self._flat_t[sel] == ti]]
                if ids.size:
                    seeds[t] = ids
            blocks = self.sampler.sample_blocks(self.g, seeds, edge_weight=self.edge_weight, key=_draw_key(self.rng),
                                                device=self.device)
            yield ({t: blocks[0].srcnodes[t].data[NID] for t in blocks[0].srctypes},
                   {t: blocks[-1].dstnodes[t].data[NID] for t in blocks[-1].dsttypes}, blocks)


class _Uniform:
    """``negative_sampler.Uniform(k)``: per positive edge (u, v) of etype (s, r, d): k edges
    (u, randint(0, num_nodes(d))), laid out k-consecutive (``src/model.py:516`` depends on that layout)."""

    def __init__(self, k):
        self.k = k

    def __call__(self, g: HeteroGraph, eids_dict, key):
        """``key``: the batch key (int) or a NumPy generator to draw one from. Destination of negative j of edge e:
        ``hash64(stream key, e * k + j) mod num_nodes`` -- the rule ``gr_negative_uniform_i64`` implements."""
        key = key if isinstance(key, int) else _draw_key(key)
        cets = g.canonical_etypes
        out = {}
        for c, eids in eids_dict.items():
            c = g.to_canonical_etype(c)
            s, _ = g.edge_arrays(c)
            e = _as_np_ids(eids).astype(np.int64)
            src = np.repeat(s[e].astype(np.int64), self.k)
            ctr = (np.repeat(e, self.k) * self.k + np.tile(np.arange(self.k, dtype=np.int64), e.size)).astype(np.uint64)
            dst = hash64(sample_key(key, NEGATIVE_STREAM + cets.index(c)), ctr) % np.uint64(g.num_nodes(c[2]))
            out[c] = (src, dst.astype(np.int64))
        return out


class negative_sampler:  # noqa: N801 - mirrors the dgl module name
    Uniform = _Uniform


class EdgeDataLoader:
    """``dgl.dataloading.EdgeDataLoader`` surface (``src/sampling.py:168-207``): iterate
    ``(input_nodes, pos_g, neg_g, blocks)``. ``pos_g`` / ``neg_g`` are compacted onto the batch's seed nodes
    (first-appearance order over [pos, neg] x etypes x (src, dst)), which are also the destination nodes of
    ``blocks[-1]``. ``exclude='reverse_types'`` drops the batch edges and their reverse-relation twins
    (same edge id) from the blocks."""

    def __init__(self, g: HeteroGraph, eids, block_sampler, g_sampling=None, exclude=None, reverse_etypes=None,
                 negative_sampler=None, batch_size=1, shuffle=False, drop_last=False, seed=0, device=None,
                 edge_weight=AUTO, **kwargs):
        _check_loader_kwargs('EdgeDataLoader', kwargs)
        self.device = device
        self.g, self.g_sampling = g, (g_sampling if g_sampling is not None else g)
        self.sampler, self.neg = block_sampler, negative_sampler
        self.edge_weight = resolve_edge_weight(self.g_sampling, edge_weight)  # blocks carry edata['occurrence'] like DGL's
        self.exclude, self.reverse_etypes = exclude, reverse_etypes or {}
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.rng = np.random.default_rng(seed)
        self.eids = {g.to_canonical_etype(c): _as_np_ids(v).astype(np.int64) for c, v in eids.items()}
        self._types = sorted(self.eids)
        self._flat_t = np.concatenate([np.full(self.eids[c].size, i) for i, c in enumerate(self._types)])
        self._flat_e = np.concatenate([self.eids[c] for c in self._types])

    def __len__(self):
        n = self._flat_e.size
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        g = self.g
        n = self._flat_e.size
        order = self.rng.permutation(n) if self.shuffle else np.arange(n)
        for b in range(len(self)):
            sel = order[b * self.batch_size:(b + 1) * self.batch_size]
            items = {c: self._flat_e[sel[self._flat_t[sel] == i]] for i, c in enumerate(self._types)}
            items = {c: e for c, e in items.items() if e.size}
            key = _draw_key(self.rng)
            if self.device is not None and torch.device(self.device).type == 'cuda':
                from .sampling_device import edge_batch_device
                yield edge_batch_device(self, items, key, torch.device(self.device))
                continue
            pos = {c: (g.edge_arrays(c)[0][e].astype(np.int64), g.edge_arrays(c)[1][e].astype(np.int64))
                   for c, e in items.items()}
            neg = self.neg(g, items, key) if self.neg is not None else {}
            space = {}
            for t in g.ntypes:
                parts = []
                for edges in (pos, neg):
                    for c in g.canonical_etypes:
                        if c in edges:
                            if c[0] == t:
                                parts.append(edges[c][0])
                            if c[2] == t:
                                parts.append(edges[c][1])
                space[t] = _first_appearance_unique(parts)
            sizes = {t: int(space[t].size) for t in g.ntypes}

            def compact(edges):
                data = {}
                for c in g.canonical_etypes:
                    s, d = edges.get(c, (np.zeros(0, np.int64), np.zeros(0, np.int64)))
                    data[c] = (_relabel(s, space[c[0]]), _relabel(d, space[c[2]]))
                cg = HeteroGraph(data, sizes)
                for t in g.ntypes:
                    cg.nodes[t].data[NID] = torch.from_numpy(space[t])
                return cg

            pos_g, neg_g = compact(pos), compact(neg)
            for c, e in items.items():
                pos_g.edges[c].data[EID] = torch.from_numpy(e)
            exclude = None
            if self.exclude == 'reverse_types':
                exclude = {}
                for c, e in items.items():
                    exclude[c] = np.concatenate([exclude.get(c, np.zeros(0, np.int64)), e])
                    rc = g.to_canonical_etype(self.reverse_etypes[c[1]])
                    exclude[rc] = np.concatenate([exclude.get(rc, np.zeros(0, np.int64)), e])
            seeds = {t: v for t, v in space.items() if v.size}
            blocks = self.sampler.sample_blocks(self.g_sampling, seeds, exclude=exclude, key=key,
                                                edge_weight=self.edge_weight)
            yield ({t: blocks[0].srcnodes[t].data[NID] for t in blocks[0].srctypes}, pos_g, neg_g, blocks)
