"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink) plumbing.

The reference is single-process; this module introduces the one strategy the path needs:

  * embedding layers -- every destination row is independent given the previous layer's rows, so each node type's
    destination id space is cut into ``world`` equal contiguous ranges; rank p runs the fused relation kernels on
    its range only (``row_begin`` / ``row_end`` of ``gr_sage_relation_f32``) and the layer output is all-gathered
    (one in-place ``all_gather_into_tensor`` per node type per layer, tables padded to ``world * chunk`` rows).
  * scoring -- the item table is sharded by the same contiguous ranges; each rank scores ALL users against its
    item shard (tcgen05 GEMM + shortlist + exact re-score = exact per-shard top-k), the per-shard lists travel to
    the rank that owns the user range (one ``all_to_all_single`` for ids, one for scores) and are merged there by
    ``gr_topk_merge`` ((score desc, id asc) -- the same total order as a single-GPU run).

No collective sits inside a kernel yet: the exchanged volumes are tiny next to the compute
(c2: 0.6 GB of embeddings, 80 MB of top-k lists per step).

The exchange helpers are backend-agnostic (``gloo`` on CPU in tests/test_distributed.py, ``nccl`` on GPUs).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def chunk_rows(n: int, world: int) -> int:
    return (n + world - 1) // world


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous id range ``[begin, end)`` owned by ``rank`` (equal chunks, the last ones may be short/empty)."""
    c = chunk_rows(n, world)
    return min(rank * c, n), min((rank + 1) * c, n)


def allgather_rows(local_full: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """``local_full``: ``[n, d]`` table in which only this rank's ``shard_range`` rows are valid. Returns the
    ``[n, d]`` table with every rank's rows filled in (view of a padded ``[world * chunk, d]`` buffer)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return local_full
    c = chunk_rows(n, world)
    b, e = shard_range(n, world, rank)
    full = local_full.new_empty((world * c, local_full.shape[1]))
    mine = full[rank * c:(rank + 1) * c]
    mine[:e - b].copy_(local_full[b:e])
    if e - b < c:
        mine[e - b:].zero_()
    if dist.get_backend(group) == 'gloo':  # gloo has no all_gather_into_tensor
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous(), group=group)
        full = torch.cat(parts, 0)
    else:
        dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n]


def balanced_bounds(block, ntype: str, world: int, row_cost: int = 8):
    """Contiguous destination ranges of ``ntype`` with (nearly) equal WORK instead of equal row counts: boundaries
    sit on the prefix sum of ``in-degree over all relations into ntype + row_cost`` (the fused kernel's cost per row
    is its gathered edges plus a fixed projection epilogue). With Zipf item popularity one hub row can hold several
    per cent of all edges, so equal row ranges leave the hub's rank far behind. Returns ``world + 1`` ints; cached on
    the block."""
    cache = block.__dict__.setdefault('_balanced_bounds', {})
    key = (ntype, world, row_cost)
    if key not in cache:
        n = block.number_of_dst_nodes(ntype)
        cost = None
        for c, rel in block.rels.items():
            if c[2] != ntype or rel.nnz == 0:
                continue
            deg = (rel.indptr[1:] - rel.indptr[:-1]).to(torch.int64)
            cost = deg if cost is None else cost + deg
        if cost is None or n == 0:
            cache[key] = [shard_range(n, world, r)[0] for r in range(world)] + [n]
        else:
            pref = torch.cumsum(cost + row_cost, 0)
            targets = (pref[-1].double() * torch.arange(1, world, dtype=torch.float64, device=pref.device) / world)
            cuts = torch.searchsorted(pref.double(), targets).clamp_(max=n).cpu().tolist()
            b = [0] + [int(x) for x in cuts] + [n]
            for i in range(1, len(b)):  # monotone (a hub heavier than a fair share empties its neighbours' ranges)
                b[i] = max(b[i], b[i - 1])
            cache[key] = b
    return cache[key]


def allgather_rows_v(local_full: torch.Tensor, bounds, group=None) -> torch.Tensor:
    """``allgather_rows`` for unequal contiguous ranges ``bounds`` (``world + 1`` ints): every rank contributes rows
    ``[bounds[rank], bounds[rank + 1])`` of its ``local_full``; chunks are padded to the longest range for one
    ``all_gather_into_tensor`` and copied back into place."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return local_full
    sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
    c = max(max(sizes), 1)
    buf = local_full.new_empty((world * c, local_full.shape[1]))
    mine = buf[rank * c:(rank + 1) * c]
    mine[:sizes[rank]].copy_(local_full[bounds[rank]:bounds[rank + 1]])
    if dist.get_backend(group) == 'gloo':
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous(), group=group)
        buf = torch.cat(parts, 0)
    else:
        dist.all_gather_into_tensor(buf, mine, group=group)
    for r in range(world):
        if r != rank and sizes[r]:
            local_full[bounds[r]:bounds[r + 1]].copy_(buf[r * c:r * c + sizes[r]])
    return local_full


def exchange_topk(ids: torch.Tensor, scores: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
    """``ids`` / ``scores``: ``[n_users, k]`` per-shard top-k of ALL users on this rank. Sends each user range to its
    owner; returns ``(ids [world, chunk, k], scores [world, chunk, k], begin, end)`` for this rank's user range
    (rows past ``end - begin`` are padding: id -1, score -inf)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n, k = ids.shape
    c = chunk_rows(n, world)
    b, e = shard_range(n, world, rank)
    if world == 1:
        return ids.unsqueeze(0), scores.unsqueeze(0), b, e
    pad = world * c - n
    if pad:
        ids = torch.cat([ids, ids.new_full((pad, k), -1)], 0)
        scores = torch.cat([scores, scores.new_full((pad, k), float('-inf'))], 0)
    dev = ids.device
    if dist.get_backend(group) == 'gloo' and ids.is_cuda:  # gloo routes through the host (tests: two ranks on one GPU)
        ids, scores = ids.cpu(), scores.cpu()
    out_ids, out_scores = torch.empty_like(ids), torch.empty_like(scores)
    dist.all_to_all_single(out_ids, ids.contiguous(), group=group)
    dist.all_to_all_single(out_scores, scores.contiguous(), group=group)
    return out_ids.view(world, c, k).to(dev), out_scores.view(world, c, k).to(dev), b, e


def sharded_get_repr(model, blocks, h: Dict[str, torch.Tensor], group=None, gather_last=None,
                     balance=()) -> Dict[str, torch.Tensor]:
    """``ConvModel.get_repr`` with destination-range sharding. Every layer output is all-gathered (the next layer
    gathers arbitrary source rows); ``gather_last`` (a collection of node types, default: all) limits the all-gather
    after the LAST layer -- a table that is not gathered comes back full-height with only this rank's rows valid
    (user-range-sharded scoring needs nothing else of the user table). Node types in ``balance`` are cut by
    ``balanced_bounds`` (equal work), the others by ``shard_range`` (equal rows -- the ranges scoring shards by)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    for i, blk in enumerate(blocks):
        bounds = {t: balanced_bounds(blk, t, world) for t in blk.dsttypes if t in balance}
        ranges = {t: ((bounds[t][rank], bounds[t][rank + 1]) if t in bounds else
                      shard_range(blk.number_of_dst_nodes(t), world, rank)) for t in blk.dsttypes}
        out = model.layers[i](blk, h, ranges)
        last = i == len(blocks) - 1
        h = {}
        for t, v in out.items():
            if last and gather_last is not None and t not in gather_last:
                h[t] = v
            elif t in bounds:
                h[t] = allgather_rows_v(v, bounds[t], group)
            else:
                h[t] = allgather_rows(v, v.shape[0], group)
    return h


def allgather_inplace(buf: torch.Tensor, world: int, rank: int, group=None) -> torch.Tensor:
    """``buf``: ``[world * chunk, d]`` whose rows ``[rank * chunk, (rank + 1) * chunk)`` this rank has filled; every other
    chunk is received in place (no staging copy: the kernels wrote straight into the collective's buffer)."""
    if world == 1:
        return buf
    c = buf.shape[0] // world
    mine = buf[rank * c:(rank + 1) * c]
    if dist.get_backend(group) == 'gloo':  # gloo has no all_gather_into_tensor
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous(), group=group)
        for r, p in enumerate(parts):
            if r != rank:
                buf[r * c:(r + 1) * c].copy_(p)
    else:
        dist.all_gather_into_tensor(buf, mine, group=group)
    return buf


def node_ranges(num_nodes: Dict[str, int], world: int, rank: int, bounds=None) -> Dict[str, Tuple[int, int]]:
    """Contiguous id ranges per node type: what a rank STORES (CSR rows, features) and computes. Equal row counts by
    default; ``bounds[t]`` (``world + 1`` ints, e.g. from ``work_bounds``) cuts type ``t`` by WORK instead."""
    bounds = bounds or {}
    return {t: ((bounds[t][rank], bounds[t][rank + 1]) if t in bounds else shard_range(n, world, rank))
            for t, n in num_nodes.items()}


def work_bounds(g, ntype: str, world: int, row_cost: int = 8):
    """``world + 1`` contiguous range boundaries of ``ntype`` with (nearly) equal aggregation WORK per range: in-degree
    over all relations into ``ntype`` plus a fixed per-row cost, computed from the host edge lists (no CSR needed, so it
    can run before a sharded ingest). With Zipf item popularity the most popular item alone holds ~8 % of all edges:
    equal ROW ranges leave its rank 1.6x over the mean at 8 ranks."""
    import numpy as np
    n = g.num_nodes(ntype)
    cost = np.full(n, row_cost, dtype=np.int64)
    for c in g.canonical_etypes:
        if c[2] == ntype:
            cost += np.bincount(g.edge_arrays(c)[1], minlength=n)
    pref = np.cumsum(cost)
    cuts = np.searchsorted(pref, pref[-1] * np.arange(1, world) / world).tolist() if n else [0] * (world - 1)
    b = [0] + [int(min(x, n)) for x in cuts] + [n]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


def sharded_forward(model, sblocks, feats_local: Dict[str, torch.Tensor], group=None, gather_last=None,
                    mark=None) -> Dict[str, torch.Tensor]:
    """Embedding pass with SHARDED STORAGE: ``sblocks`` are this rank's ``HeteroGraph.sharded_block_on`` blocks (one per
    conv layer; equal ``shard_range`` rows), ``feats_local[t]`` the raw feature rows of this rank's id range of node type
    ``t``. Each rank holds only its own feature rows; they are all-gathered (raw, 8 - 16 bytes per row -- or embedded,
    when the embedding is not wider than the features) so that every rank has the embedded inputs of ALL source rows
    (the next layer gathers arbitrary source rows), and every conv layer computes this rank's destination rows
    straight into the all-gather buffer. Per-rank resident graph + features are ~1/world of the whole; the gathered ``[N, D]`` tables
    are not (every rank reads all source rows). ``gather_last``: node types to all-gather after the LAST layer
    (default all); a type left out comes back full-height with only this rank's rows valid. ``mark(name)`` (optional)
    is called after the input embedding + its gather ('embed_in'), after every layer's kernels ('compute<i>') and after
    every layer's all-gather ('gather<i>') -- bench.py separates kernel time from collective time with it.
    ``relu(fc_preagg(h))`` of the ``*_nn`` aggregators runs on all source rows on every rank: gathering the projected
    table instead would move more bytes over NVLink than the (tensor-core) projection costs."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = sblocks[0].shard_ranges
    vbounds = getattr(sblocks[0], 'shard_bounds', None) or {}   # node types cut by work: unequal chunks
    num = sblocks[0].num_src
    mark = mark or (lambda name: None)

    def gather_rows(own: torch.Tensor, t: str) -> torch.Tensor:
        """Rows [b, e) of node type t (this rank's) -> the full [num[t], d] table on every rank."""
        b, e = ranges[t]
        if t in vbounds:
            full = own.new_empty((num[t], own.shape[1]))
            full[b:e].copy_(own)
            return allgather_rows_v(full, vbounds[t], group)
        c = chunk_rows(num[t], world)
        buf = own.new_empty((world * c, own.shape[1]))
        buf[rank * c:rank * c + (e - b)].copy_(own)
        if e - b < c:
            buf[rank * c + (e - b):(rank + 1) * c].zero_()
        return allgather_inplace(buf, world, rank, group)[:num[t]]

    h = {}
    for t, x in feats_local.items():
        d_embed = getattr(model, t + '_embed').proj_feats.out_features
        if x.shape[1] < d_embed:
            # the raw feature rows (a few columns) are far smaller than the embedded ones: all-gather THOSE and embed
            # every row locally -- NodeEmbedding is a pure output-write stream (c2: 0.15 ms for all rows), the gather
            # of the embedded table would move 512 bytes per row over NVLink instead of 8 - 16
            h[t] = model.embed_type(t, gather_rows(x, t))
        else:
            h[t] = gather_rows(model.embed_type(t, x), t)
    mark('embed_in')
    for i, blk in enumerate(sblocks):
        last = i == len(sblocks) - 1
        layer = model.layers[i]
        d_out = next(iter(layer.mods.values()))._out_feats
        ref = h[next(iter(h))]
        bufs = {}
        for t in blk.dsttypes:
            if t in vbounds:   # unequal chunks: the kernels write rows [b, e) of a full-height table
                bufs[t] = ref.new_empty((num[t], d_out))
                continue
            c = chunk_rows(num[t], world)
            bufs[t] = ref.new_empty((world * c, d_out))
            b, e = ranges[t]
            if e - b < c:
                bufs[t][rank * c + (e - b):(rank + 1) * c].zero_()
        out = layer(blk, h, out_buffers=bufs)
        mark('compute%d' % i)
        h = {}
        for t, v in out.items():
            if not (last and gather_last is not None and t not in gather_last):
                if t in vbounds:
                    allgather_rows_v(bufs[t], vbounds[t], group)
                else:
                    allgather_inplace(bufs[t], world, rank, group)
            h[t] = bufs[t][:num[t]]
        mark('gather%d' % i)
    return h


def choose_item_shards(n_users: int, n_items: int, world: int) -> int:
    """Scoring layout: ``world`` = item-range shards + owner-side merge (every rank preps / re-scores ALL users);
    ``1`` = user-range shards against a replicated item table (every rank preps ALL items, no exchange).
    The repeated per-row work is what differs, so shard the longer side."""
    return 1 if n_users >= n_items else world


def sharded_recommend(h_user: torch.Tensor, h_item: torch.Tensor, k: int, bought=None, config=None, group=None,
                      mark=None, item_shards: Optional[int] = None, return_overflow: bool = False):
    """Sharded scoring. ``h_user`` / ``h_item`` are the full tables (every rank holds them after the last
    all-gather); ``bought`` rows follow ``h_user`` rows. Returns ``(ids [u_loc, k], scores, (begin, end))`` for the
    user range this rank owns.

    ``item_shards = world``: the item table is cut into ``world`` contiguous id ranges, every rank scores ALL
    users against its range and the per-shard exact top-k lists are merged on the rank owning the user range.
    ``item_shards = 1``: every rank scores only its own contiguous user range against the whole (replicated) item
    table -- no exchange, and the per-user prep / re-score work is not repeated on every rank.
    ``item_shards = None`` (default) picks by ``choose_item_shards`` (shard the longer side).
    Only the rows a layout reads need to be valid: the user-range layout reads this rank's user rows and ALL item rows,
    the item-range layout ALL user rows and this rank's item rows (``sharded_forward(gather_last=...)``)."""
    from . import ops
    from .recs import RecsConfig, ScoringTable, recommend_topk
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cfg = config or RecsConfig()
    if item_shards is None:
        item_shards = choose_item_shards(h_user.shape[0], h_item.shape[0], world)
    if item_shards == 1 and world > 1:
        ub, ue = shard_range(h_user.shape[0], world, rank)
        table = ScoringTable(h_item, cfg)
        sub = None
        if bought is not None:  # the row slice of this rank, cached on the parent CSR (host slicing + H2D once)
            cache = bought.__dict__.setdefault('_range_cache', {})
            if (ub, ue) not in cache:
                cache[(ub, ue)] = bought.select(range(ub, ue))
            sub = cache[(ub, ue)]
        ids, scores, n_over = recommend_topk(h_user[ub:ue], table, k, sub, mark=mark, return_overflow=True)
        return (ids, scores, (ub, ue), n_over) if return_overflow else (ids, scores, (ub, ue))
    if item_shards != world:
        raise ValueError('item_shards must be 1 or the world size')
    ib, ie = shard_range(h_item.shape[0], world, rank)
    table = ScoringTable(h_item[ib:ie], cfg, item_id_base=ib)
    ids, scores, n_over = recommend_topk(h_user, table, k, bought, mark=mark, return_overflow=True)
    all_ids, all_scores, ub, ue = exchange_topk(ids, scores, group)
    if world > 1:
        m_scores, m_ids = ops.topk_merge(all_scores.contiguous(), all_ids.contiguous(), k)
        ids, scores = m_ids[:ue - ub], m_scores[:ue - ub]
    return (ids, scores, (ub, ue), n_over) if return_overflow else (ids, scores, (ub, ue))
