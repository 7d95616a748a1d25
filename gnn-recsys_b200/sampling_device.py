"""Sampled blocks built on the device -- the ``device=`` path of the loaders in
``dataloading.py``.

What ``dgl.dataloading`` does on CPU worker processes behind the reference's loaders (``src/sampling.py:153-207``,
consumed by ``src/train/run.py:89-139``) runs here as CUDA kernels on the graph structure that already sits in HBM
(``HeteroGraph.full_block_on``: per relation an int32 CSR over destination rows + edge permutation):

  frontier of a layer    ``gr_sample_count_i32`` + ``gr_sample_fill_i32`` per relation (fan-out without replacement by
                         counter-based tickets, or the full neighbourhood; the batch's own edges and their reverse
                         twins excluded)
  ``to_block``           ``gr_remap_first_appearance_i64`` over ``[seeds | frontier sources ...]`` per source node type:
                         seeds keep ids ``0..n_dst-1`` (the block prefix invariant), new sources follow in CSR order
  negatives              ``gr_negative_uniform_i64`` (k consecutive per positive edge)
  pos_g / neg_g          the same remap over ``[pos, neg] x etypes x (src, dst)`` -> batch-local int32 end points

PyTorch is plumbing: buffers, ``torch.cat`` / row gathers of features, and one host read per layer of the frontier
sizes (the blocks have data-dependent shapes). The host builder in ``dataloading.py`` applies the same rules with the
same keys and yields identical blocks (tests/test_gpu_parity.py::test_device_blocks_*).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from .graph import Block, DeviceEdgeGraph, HeteroGraph, Relation, NID, EID, _as_np_ids, INT32_MAX
from .dataloading import sample_key, NEGATIVE_STREAM


def _ids_on(device, v) -> torch.Tensor:
    if isinstance(v, torch.Tensor):
        return v.to(device=device, dtype=torch.int64).reshape(-1).contiguous()
    return torch.from_numpy(_as_np_ids(v).astype(np.int64)).to(device)


def _exclusion_lists(exclude, device) -> Dict:
    """Per canonical etype: ascending int32 edge ids on the device (``excl_sorted`` of ``gr_sample_*``)."""
    out = {}
    for c, v in (exclude or {}).items():
        v = _ids_on(device, v)
        if v.numel():
            out[c] = torch.sort(v).values.to(torch.int32)
    return out


def sample_blocks_device(g: HeteroGraph, sampler, seed_nodes, key: int, device: torch.device, exclude=None,
                         edge_weight: Optional[str] = None) -> List[Block]:
    full = g.full_block_on(device, edge_weight)
    cets = g.canonical_etypes
    excl = _exclusion_lists(exclude, device)
    seeds = {t: _ids_on(device, v) for t, v in seed_nodes.items()}
    blocks: List[Block] = []
    for layer in reversed(range(sampler.num_layers)):
        fan = sampler._fanout(layer) or 0
        counted = {}
        for ci, c in enumerate(cets):
            rel, sd = full.rels[c], seeds.get(c[2])
            if sd is None or sd.numel() == 0 or rel.nnz == 0:
                continue
            out_indptr, total = ops.sample_count(rel.indptr, rel.eperm, sd, fan, excl.get(c))
            counted[c] = (ci, out_indptr, total)
        totals = torch.cat([v[2] for v in counted.values()]).tolist() if counted else []  # the layer's one host read
        totals = dict(zip(counted.keys(), totals))
        rels, src_ids, num_src = {}, {}, {}
        local_of, bufs = {}, []
        for t in g.ntypes:
            mine = [c for c in cets if c[0] == t and c in counted]
            n_seed = int(seeds[t].numel()) if t in seeds else 0
            n_all = n_seed + sum(totals[c] for c in mine)
            if n_all > INT32_MAX:
                raise OverflowError('frontier of node type %r exceeds int32' % t)
            buf = torch.empty(n_all, dtype=torch.int64, device=device)
            if n_seed:
                buf[:n_seed] = seeds[t]
            off = n_seed
            for c in mine:
                ci, out_indptr, _ = counted[c]
                n = totals[c]
                eid = torch.empty(n, dtype=torch.int32, device=device)
                if n:
                    rel = full.rels[c]
                    ops.sample_fill(rel.indptr, rel.indices, rel.eperm, seeds[c[2]], fan, excl.get(c),
                                    sample_key(key, layer * 64 + ci), out_indptr, buf[off:off + n], eid)
                local_of[c] = (off, n, eid)
                off += n
            bufs.append(buf)
        for t, (new_ids, uniq) in zip(g.ntypes, ops.remap_many(bufs)):  # to_block; one host read for all node types
            src_ids[t], num_src[t] = uniq, int(uniq.numel())
            for c in cets:
                if c[0] != t or c not in counted:
                    continue
                off, n, eid = local_of[c]
                w = None
                if full.rels[c].weight is not None:
                    w = _edge_weight_by_eid(g, c, edge_weight, device)[eid.long()]
                rels[c] = Relation(counted[c][1], new_ids[off:off + n], num_src[t],
                                   int(seeds[c[2]].numel()), eid, w)
        for c in cets:
            if c not in rels:
                n_dst = int(seeds[c[2]].numel()) if c[2] in seeds else 0
                w = torch.zeros(0, dtype=torch.float32, device=device) if full.rels[c].weight is not None else None
                rels[c] = Relation(torch.zeros(n_dst + 1, dtype=torch.int32, device=device),
                                   torch.zeros(0, dtype=torch.int32, device=device), num_src[c[0]], n_dst,
                                   torch.zeros(0, dtype=torch.int32, device=device), w)
        num_dst = {t: int(seeds[t].numel()) if t in seeds else 0 for t in g.ntypes}
        sf, df = {}, {}
        for t in g.ntypes:
            data = g.device_node_data(t, device)
            did = seeds[t] if t in seeds else torch.zeros(0, dtype=torch.int64, device=device)
            sf[t] = {k: v[src_ids[t]] for k, v in data.items()}
            sf[t][NID] = src_ids[t]
            df[t] = {k: v[did] for k, v in data.items()}
            df[t][NID] = did
        blocks.insert(0, Block(rels, num_src, num_dst, sf, df))
        seeds = dict(src_ids)
    return blocks


def _edge_weight_by_eid(g: HeteroGraph, c, name: str, device) -> torch.Tensor:
    """Per-edge scalar of one relation indexed by EDGE ID on the device (cached with the graph)."""
    key = ('w', c, name, str(device))
    if key not in g._dev_edges:
        g._dev_edges[key] = g.edges[c].data[name].to(device).to(torch.float32).reshape(-1).contiguous()
    return g._dev_edges[key]


def edge_batch_device(loader, items: Dict, key: int, device: torch.device):
    """One ``EdgeDataLoader`` batch on the device: ``(input_nodes, pos_g, neg_g, blocks)``; ``items`` = the batch's
    edge ids per canonical etype (host arrays, a few KB -- the only thing that crosses PCIe)."""
    g = loader.g
    cets = g.canonical_etypes
    eids = {c: torch.from_numpy(e.astype(np.int64)).to(device) for c, e in items.items()}
    pos, neg = {}, {}
    for c, e in eids.items():
        u_all, v_all = g.device_edges(c, device)
        pos[c] = (u_all[e].long(), v_all[e].long())
        if loader.neg is not None:
            neg[c] = ops.negative_uniform(u_all, e, loader.neg.k, g.num_nodes(c[2]),
                                          sample_key(key, NEGATIVE_STREAM + cets.index(c)))
    # batch node space per type: first appearance over [pos, neg] x etypes x (src, dst)
    space, local, cats, layout = {}, {}, [], []
    for t in g.ntypes:
        parts, slots = [], []
        for name, edges in (('pos', pos), ('neg', neg)):
            for c in cets:
                if c in edges:
                    if c[0] == t:
                        parts.append(edges[c][0])
                        slots.append((name, c, 0))
                    if c[2] == t:
                        parts.append(edges[c][1])
                        slots.append((name, c, 1))
        cats.append(torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64, device=device))
        layout.append((parts, slots))
    for t, (parts, slots), (new_ids, uniq) in zip(g.ntypes, layout, ops.remap_many(cats)):
        space[t] = uniq
        off = 0
        for part, slot in zip(parts, slots):
            local[slot] = new_ids[off:off + part.numel()]
            off += part.numel()
    sizes = {t: int(space[t].numel()) for t in g.ntypes}
    empty = torch.zeros(0, dtype=torch.int32, device=device)

    def compact(name, edges):
        return DeviceEdgeGraph({c: ((local[(name, c, 0)], local[(name, c, 1)]) if c in edges else (empty, empty))
                                for c in cets}, sizes, space,
                               {c: eids[c] for c in edges} if name == 'pos' else None)

    pos_g, neg_g = compact('pos', pos), compact('neg', neg)
    exclude = None
    if loader.exclude == 'reverse_types':
        exclude = {}
        for c, e in eids.items():
            exclude[c] = torch.cat([exclude[c], e]) if c in exclude else e
            rc = g.to_canonical_etype(loader.reverse_etypes[c[1]])
            exclude[rc] = torch.cat([exclude[rc], e]) if rc in exclude else e
    seeds = {t: v for t, v in space.items() if v.numel()}
    blocks = sample_blocks_device(loader.g_sampling, loader.sampler, seeds, key, device, exclude, loader.edge_weight)
    return ({t: blocks[0].srcnodes[t].data[NID] for t in blocks[0].srctypes}, pos_g, neg_g, blocks)
