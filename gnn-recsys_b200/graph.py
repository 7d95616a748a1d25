"""Host-side graph containers for the B200 hot path.

The reference keeps its user-item graph in a ``dgl.DGLHeteroGraph`` (built by
``src/builder.py:377-383`` from the ``graph_schema`` dict of ``src/utils_data.py:204-238``) and feeds the
model with DGL *blocks* (``main_inference.py:126-138``, ``src/train/run.py:334-346``). This module is the
native replacement for that object surface: a ``HeteroGraph`` that answers the
queries the hot path makes (``ntypes``, ``canonical_etypes``, ``num_nodes``, ``nodes[nt].data``,
``find_edges``, ``out_edges``, ``all_edges`` ...) and a ``Block`` that stores, per relation, an **int32
CSR over destination rows** -- the layout the CUDA gather-reduce kernel reads.

Layout rules (same as ``dgl.heterograph``): node types sorted, canonical etypes sorted as tuples,
edge id = position in the input list, ``num_nodes`` = max id + 1 unless given explicitly.
CSR construction is a *stable* sort of the edge list by destination id, so neighbours of a row stay
in edge-id order; it is bit-exact against the CPU oracle's restatement (tests/test_host.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

NID = '_ID'
EID = '_ID'
INT32_MAX = 2 ** 31 - 1

CEType = Tuple[str, str, str]


def _as_np_ids(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x).reshape(-1))


def csr_by_dst_host(src: np.ndarray, dst: np.ndarray, n_dst: int):
    """Stable COO -> CSR-over-destination on the host (int32 out, overflow-guarded).

    Returns ``(indptr[n_dst+1], indices[nnz], eperm[nnz])`` where ``eperm[j]`` is the edge id stored
    in CSR slot ``j``. Replaces DGL's internal CSC construction behind ``update_all``
    (``src/model.py:145-147``); the device version is ``gr_csr_build_i32``.
    """
    src = _as_np_ids(src)
    dst = _as_np_ids(dst)
    nnz = int(src.shape[0])
    if nnz > INT32_MAX or n_dst > INT32_MAX:
        raise OverflowError('int32 CSR cannot index %d edges / %d rows' % (nnz, n_dst))
    if nnz and (int(dst.max()) >= n_dst or int(dst.min()) < 0):
        raise IndexError('destination id out of range')
    eperm = np.argsort(dst, kind='stable').astype(np.int32)
    indices = src[eperm].astype(np.int32)
    counts = np.bincount(dst, minlength=n_dst) if nnz else np.zeros(n_dst, dtype=np.int64)
    indptr = np.zeros(n_dst + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr.astype(np.int32), indices, eperm


@dataclass
class Relation:
    """One canonical etype of a ``Block`` as CSR over destination rows (all tensors share a device)."""
    indptr: torch.Tensor                     # int32 [n_dst + 1]
    indices: torch.Tensor                    # int32 [nnz], source row ids (block-local)
    n_src: int
    n_dst: int
    eperm: Optional[torch.Tensor] = None     # int32 [nnz], edge id held by each CSR slot
    weight: Optional[torch.Tensor] = None    # float32 [nnz], per-edge scalar in CSR order (``*_edge`` aggregators)

    @property
    def nnz(self) -> int:
        return int(self.indices.shape[0])

    def to(self, device, non_blocking=False) -> 'Relation':
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)  # noqa: E731
        return Relation(mv(self.indptr), mv(self.indices), self.n_src, self.n_dst, mv(self.eperm), mv(self.weight))

    def pin(self) -> 'Relation':
        pn = lambda t: None if t is None else t.pin_memory()  # noqa: E731
        return Relation(pn(self.indptr), pn(self.indices), self.n_src, self.n_dst, pn(self.eperm), pn(self.weight))

    def csr_bytes(self) -> int:
        return 4 * (self.n_dst + 1) + 4 * self.nnz


class _Frames:
    """Typed frame view: ``view[key]`` -> tensor (one type) or dict type -> tensor (several types)."""

    def __init__(self, frames: Dict):
        self._frames = frames

    def __getitem__(self, key):
        if len(self._frames) == 1:
            return next(iter(self._frames.values()))[key]
        return {t: f[key] for t, f in self._frames.items() if key in f}

    def __setitem__(self, key, val):
        if len(self._frames) == 1 and not isinstance(val, dict):
            next(iter(self._frames.values()))[key] = val
            return
        for t, v in val.items():
            self._frames[t][key] = v

    def __contains__(self, key):
        return any(key in f for f in self._frames.values())


class _Typed:
    def __init__(self, frame):
        self.data = frame


class _TypedIndex:
    def __init__(self, frames, canon=None):
        self._frames, self._canon = frames, canon

    def __getitem__(self, key):
        if self._canon is not None:
            key = self._canon(key)
        return _Typed(self._frames[key])


class _EdgeIndex(_TypedIndex):
    def __init__(self, g):
        super().__init__(g._edge_frames, g.to_canonical_etype)
        self._g = g

    def __call__(self, etype=None, form='uv'):
        return self._g.all_edges(form=form, etype=etype)


class Block:
    """A bipartite message-passing block (DGL block surface used by ``src/model.py`` / ``run.py:334-346``).

    ``num_src[nt]`` source rows and ``num_dst[nt]`` destination rows per node type; destination nodes
    are the first ``num_dst[nt]`` source nodes (the DGL block prefix invariant the reference's
    ``HeteroGraphConv`` relies on). For a full-graph block ``num_src == num_dst == num_nodes``.
    """
    is_block = True

    def __init__(self, rels: Dict[CEType, Relation], num_src: Dict[str, int], num_dst: Dict[str, int],
                 src_frames=None, dst_frames=None):
        self.rels = dict(sorted(rels.items()))
        self.num_src = dict(num_src)
        self.num_dst = dict(num_dst)
        self._src_frames = src_frames if src_frames is not None else {t: {} for t in sorted(num_src)}
        self._dst_frames = dst_frames if dst_frames is not None else {t: {} for t in sorted(num_dst)}

    @property
    def canonical_etypes(self):
        return list(self.rels.keys())

    @property
    def srctypes(self):
        return sorted(self.num_src)

    @property
    def dsttypes(self):
        return sorted(self.num_dst)

    ntypes = srctypes

    @property
    def srcdata(self):
        return _Frames(self._src_frames)

    @property
    def dstdata(self):
        return _Frames(self._dst_frames)

    @property
    def srcnodes(self):
        return _TypedIndex(self._src_frames)

    @property
    def dstnodes(self):
        return _TypedIndex(self._dst_frames)

    def number_of_src_nodes(self, ntype):
        return self.num_src[ntype]

    def number_of_dst_nodes(self, ntype):
        return self.num_dst.get(ntype, 0)

    def number_of_edges(self, etype=None):
        if etype is None:
            return sum(r.nnz for r in self.rels.values())
        if not isinstance(etype, tuple):
            etype = [c for c in self.rels if c[1] == etype][0]
        return self.rels[etype].nnz

    num_edges = number_of_edges

    @property
    def device(self):
        for r in self.rels.values():
            return r.indptr.device
        return torch.device('cpu')

    def to(self, device, non_blocking=False) -> 'Block':
        mvf = lambda fr: {t: {k: v.to(device, non_blocking=non_blocking) for k, v in f.items()}  # noqa: E731
                          for t, f in fr.items()}
        return Block({c: r.to(device, non_blocking) for c, r in self.rels.items()}, self.num_src, self.num_dst,
                     mvf(self._src_frames), mvf(self._dst_frames))

    def pin(self) -> 'Block':
        pnf = lambda fr: {t: {k: v.pin_memory() for k, v in f.items()} for t, f in fr.items()}  # noqa: E731
        return Block({c: r.pin() for c, r in self.rels.items()}, self.num_src, self.num_dst,
                     pnf(self._src_frames), pnf(self._dst_frames))

    def csr_bytes(self) -> int:
        return sum(r.csr_bytes() for r in self.rels.values())


class HeteroGraph:
    """COO heterograph with lazily built per-relation CSR (``dgl.heterograph`` surface of SURVEY.md 8b)."""
    is_block = False

    def __init__(self, data_dict, num_nodes_dict: Optional[Dict[str, int]] = None):
        edges = {}
        for c, data in data_dict.items():
            if isinstance(data, tuple) and len(data) == 2 and not np.isscalar(data[0]):
                s, d = _as_np_ids(data[0]), _as_np_ids(data[1])
            else:  # list of (src, dst) tuples, as built by src/utils_data.py:204-214
                arr = np.asarray(list(data), dtype=np.int64).reshape(-1, 2)
                s, d = np.ascontiguousarray(arr[:, 0]), np.ascontiguousarray(arr[:, 1])
            if s.shape != d.shape:
                raise ValueError('src/dst length mismatch for %r' % (c,))
            edges[tuple(c)] = (s, d)
        self._edges = dict(sorted(edges.items()))
        num = {}
        for (st, rn, dt), (s, d) in self._edges.items():
            if s.size and (int(s.min()) < 0 or int(d.min()) < 0):  # DGL rejects these too; the int32 device CSR would wrap
                raise ValueError('negative node id in relation %r' % ((st, rn, dt),))
            num[st] = max(num.get(st, 0), int(s.max()) + 1 if s.size else 0)
            num[dt] = max(num.get(dt, 0), int(d.max()) + 1 if d.size else 0)
        if num_nodes_dict:
            for t, n in num_nodes_dict.items():
                if n < num.get(t, 0):
                    raise ValueError('num_nodes[%s]=%d smaller than max id + 1' % (t, n))
                num[t] = int(n)
        self._num = dict(sorted(num.items()))
        self._node_frames = {t: {} for t in self._num}
        self._edge_frames = {c: {} for c in self._edges}
        self._csr_cache: Dict[CEType, Tuple[np.ndarray, np.ndarray, np.ndarray]] = {}
        self._dev_blocks: Dict = {}
        self._dev_edges: Dict = {}
        self._dev_node_data: Dict = {}

    # ---- metagraph ----
    @property
    def ntypes(self):
        return list(self._num.keys())

    @property
    def canonical_etypes(self):
        return list(self._edges.keys())

    @property
    def etypes(self):
        return [c[1] for c in self._edges]

    def to_canonical_etype(self, etype) -> CEType:
        if etype is None:
            if len(self._edges) != 1:
                raise KeyError('etype required on a multi-relation graph')
            return next(iter(self._edges))
        if isinstance(etype, tuple):
            if etype not in self._edges:
                raise KeyError(etype)
            return etype
        hits = [c for c in self._edges if c[1] == etype]
        if len(hits) != 1:
            raise KeyError(etype)
        return hits[0]

    # ---- sizes ----
    def num_nodes(self, ntype=None):
        return sum(self._num.values()) if ntype is None else self._num[ntype]

    number_of_nodes = num_nodes

    def num_edges(self, etype=None):
        if etype is None:
            return sum(int(e[0].shape[0]) for e in self._edges.values())
        return int(self._edges[self.to_canonical_etype(etype)][0].shape[0])

    number_of_edges = num_edges

    # ---- frames ----
    @property
    def nodes(self):
        return _TypedIndex(self._node_frames)

    @property
    def edges(self):
        return _EdgeIndex(self)

    @property
    def ndata(self):
        return _Frames(self._node_frames)

    @property
    def edata(self):
        return _Frames(self._edge_frames)

    # ---- structure queries (ids come back as int64 tensors, like DGL) ----
    def all_edges(self, form='uv', order=None, etype=None):
        s, d = self._edges[self.to_canonical_etype(etype)]
        if form == 'eid':
            return torch.arange(s.shape[0])
        u, v = torch.from_numpy(s.astype(np.int64)), torch.from_numpy(d.astype(np.int64))
        return (u, v) if form == 'uv' else (u, v, torch.arange(s.shape[0]))

    def find_edges(self, eid, etype=None):
        s, d = self._edges[self.to_canonical_etype(etype)]
        eid = _as_np_ids(eid).astype(np.int64)
        return torch.from_numpy(s[eid].astype(np.int64)), torch.from_numpy(d[eid].astype(np.int64))

    def out_edges(self, u, form='uv', etype=None):
        c = self.to_canonical_etype(etype)
        s, d = self._edges[c]
        mask = np.zeros(self._num[c[0]], dtype=bool)
        mask[_as_np_ids(u).astype(np.int64)] = True
        eid = np.nonzero(mask[s])[0]
        if form == 'eid':
            return torch.from_numpy(eid)
        uu, vv = torch.from_numpy(s[eid].astype(np.int64)), torch.from_numpy(d[eid].astype(np.int64))
        return (uu, vv) if form == 'uv' else (uu, vv, torch.from_numpy(eid))

    def edge_arrays(self, etype):
        """Raw host (src, dst) numpy arrays of one relation (no copy)."""
        return self._edges[self.to_canonical_etype(etype)]

    # ---- CSR / blocks ----
    def csr(self, etype):
        c = self.to_canonical_etype(etype)
        if c not in self._csr_cache:
            s, d = self._edges[c]
            self._csr_cache[c] = csr_by_dst_host(s, d, self._num[c[2]])
        return self._csr_cache[c]

    def full_block(self, edge_weight: Optional[str] = None, with_features: bool = True) -> Block:
        """One block over the whole graph (identity relabel): what a full-neighbour sampler seeded with
        *all* nodes produces (``main_inference.py:126-138`` without the 128-node batching)."""
        rels = {}
        for c in self._edges:
            indptr, indices, eperm = self.csr(c)
            w = None
            if edge_weight is not None and edge_weight in self._edge_frames[c]:
                w = self._edge_frames[c][edge_weight].detach().cpu().to(torch.float32).reshape(-1)[
                    torch.from_numpy(eperm.astype(np.int64))].contiguous()
            rels[c] = Relation(torch.from_numpy(indptr), torch.from_numpy(indices), self._num[c[0]], self._num[c[2]],
                               torch.from_numpy(eperm), w)
        frames = {t: (dict(f) if with_features else {}) for t, f in self._node_frames.items()}
        ids = {t: torch.arange(n) for t, n in self._num.items()}
        sf = {t: dict(frames[t], **{NID: ids[t]}) for t in self._num}
        df = {t: dict(frames[t], **{NID: ids[t]}) for t in self._num}
        return Block(rels, self._num, self._num, sf, df)

    def full_block_on(self, device, edge_weight: Optional[str] = None) -> Block:
        """Device-resident full-graph block, built once per (device, edge weight) and kept: the graph structure
        stays in HBM across ``get_embeddings`` calls instead of being re-sent per batch (``run.py:338-339``).
        The COO lists go to the device as int32 and the stable CSR is built there (``gr_csr_build_i32``).
        Node features are NOT cached here -- they travel with each call."""
        key = (str(device), edge_weight)
        if key not in self._dev_blocks:
            from . import ops
            rels = {}
            for c, (s, d) in self._edges.items():
                if s.shape[0] > INT32_MAX or max(self._num[c[0]], self._num[c[2]]) > INT32_MAX:
                    raise OverflowError('int32 CSR cannot index relation %r' % (c,))
                src = torch.from_numpy(s.astype(np.int32, copy=False)).to(device)
                dst = torch.from_numpy(d.astype(np.int32, copy=False)).to(device)
                indptr, indices, eperm = ops.csr_build(src, dst, self._num[c[2]])
                w = None
                if edge_weight is not None and edge_weight in self._edge_frames[c]:
                    w = self._edge_frames[c][edge_weight].to(device).to(torch.float32).reshape(-1)[eperm.long()].contiguous()
                rels[c] = Relation(indptr, indices, self._num[c[0]], self._num[c[2]], eperm, w)
            self._dev_blocks[key] = Block(rels, self._num, self._num)
        return self._dev_blocks[key]

    def sharded_block_on(self, device, ranges: Dict[str, Tuple[int, int]], edge_weight: Optional[str] = None,
                         bounds=None) -> Block:
        """This rank's SHARD of the full-graph block (multi-GPU, ``distributed.py``): for every relation only the CSR
        rows of the destination range ``ranges[dst ntype] = (begin, end)`` -- ``indptr`` has ``end - begin + 1``
        entries and starts at 0, ``indices`` are GLOBAL source ids, ``eperm`` holds the global edge id of each slot.
        Only this rank's edges travel to the device, so the resident graph is ~1/world of the full CSR
        (``Block.shard_ranges`` marks the block; ``HeteroGraphConv`` then computes exactly those rows). ``bounds``:
        ``{ntype: world + 1 boundaries}`` of the node types whose ranges are NOT the equal-row ``shard_range`` cut
        (``distributed.work_bounds``), recorded so that ``sharded_forward`` all-gathers them as unequal chunks."""
        key = (str(device), edge_weight, tuple(sorted(ranges.items())))
        if key not in self._dev_blocks:
            from . import ops
            rels = {}
            for c, (s, d) in self._edges.items():
                if s.shape[0] > INT32_MAX or max(self._num[c[0]], self._num[c[2]]) > INT32_MAX:
                    raise OverflowError('int32 CSR cannot index relation %r' % (c,))
                b, e = ranges[c[2]]
                sel = np.nonzero((d >= b) & (d < e))[0]            # ascending edge ids: the stable order survives
                src = torch.from_numpy(s[sel].astype(np.int32, copy=False)).to(device)
                dst = torch.from_numpy((d[sel] - b).astype(np.int32, copy=False)).to(device)
                indptr, indices, eperm = ops.csr_build(src, dst, e - b)
                eid = torch.from_numpy(sel.astype(np.int64)).to(device)[eperm.long()]
                w = None
                if edge_weight is not None and edge_weight in self._edge_frames[c]:
                    w = self._edge_frames[c][edge_weight].reshape(-1)[eid.cpu()].to(device).to(torch.float32).contiguous()
                rels[c] = Relation(indptr, indices, self._num[c[0]], e - b, eid.to(torch.int32), w)
            blk = Block(rels, self._num, self._num)
            blk.shard_ranges = {t: (int(b), int(e)) for t, (b, e) in ranges.items()}
            blk.shard_bounds = {t: [int(x) for x in v] for t, v in (bounds or {}).items()}  # types cut by work (unequal)
            self._dev_blocks[key] = blk
        return self._dev_blocks[key]

    def device_edges(self, etype, device):
        """(src, dst) of one relation as int32 device tensors (cached) -- what the edge-scoring kernel reads."""
        c = self.to_canonical_etype(etype)
        key = (c, str(device))
        if key not in self._dev_edges:
            s, d = self._edges[c]
            if s.size and (int(s.max()) > INT32_MAX or int(d.max()) > INT32_MAX):
                raise OverflowError('node ids exceed int32')
            self._dev_edges[key] = (torch.from_numpy(s.astype(np.int32)).to(device),
                                    torch.from_numpy(d.astype(np.int32)).to(device))
        return self._dev_edges[key]

    def device_node_data(self, ntype, device) -> Dict[str, torch.Tensor]:
        """Node frame of one type as device tensors (cached per host tensor) -- device-built blocks gather their
        ``srcdata`` / ``dstdata`` rows from it instead of re-sending features with every batch."""
        out = {}
        for k, v in self._node_frames[ntype].items():
            key = (ntype, str(device), k)
            hit = self._dev_node_data.get(key)
            if hit is None or hit[0] is not v or hit[1] != v._version:
                hit = (v, v._version, v.to(device))
                self._dev_node_data[key] = hit
            out[k] = hit[2]
        return out

    def to(self, device, **kwargs):
        return self  # structure stays on the host; blocks carry the device-resident CSR


def heterograph(data_dict, num_nodes_dict=None) -> HeteroGraph:
    """``dgl.heterograph`` replacement (``src/builder.py:382``)."""
    return HeteroGraph(data_dict, num_nodes_dict)


def edge_graph(parent_ntype_sizes: Dict[str, int], edges: Dict[CEType, Tuple]) -> HeteroGraph:
    """Graph holding only edges to be scored (the ``pos_g`` / ``neg_g`` of ``src/model.py:423-470``)."""
    return HeteroGraph(edges, parent_ntype_sizes)


class DeviceEdgeGraph:
    """Edges to be scored, already on the device: the ``pos_g`` / ``neg_g`` a device-side ``EdgeDataLoader`` yields
    (``src/train/run.py:104-116``). Node ids are local to the batch's seed space (= destination nodes of
    ``blocks[-1]``); ``nodes[nt].data[NID]`` holds the global ids, ``edges[c].data[EID]`` the parent edge ids."""
    is_block = False

    def __init__(self, edges: Dict[CEType, Tuple[torch.Tensor, torch.Tensor]], num_nodes: Dict[str, int],
                 node_ids: Optional[Dict[str, torch.Tensor]] = None, edge_ids: Optional[Dict] = None):
        self._edges = dict(sorted(edges.items()))
        self._num = dict(sorted(num_nodes.items()))
        self._node_frames = {t: ({NID: node_ids[t]} if node_ids and t in node_ids else {}) for t in self._num}
        self._edge_frames = {c: ({EID: edge_ids[c]} if edge_ids and c in edge_ids else {}) for c in self._edges}

    @property
    def ntypes(self):
        return list(self._num.keys())

    @property
    def canonical_etypes(self):
        return list(self._edges.keys())

    @property
    def etypes(self):
        return [c[1] for c in self._edges]

    to_canonical_etype = HeteroGraph.to_canonical_etype

    def num_nodes(self, ntype=None):
        return sum(self._num.values()) if ntype is None else self._num[ntype]

    number_of_nodes = num_nodes

    def num_edges(self, etype=None):
        if etype is None:
            return sum(int(e[0].shape[0]) for e in self._edges.values())
        return int(self._edges[self.to_canonical_etype(etype)][0].shape[0])

    number_of_edges = num_edges

    @property
    def nodes(self):
        return _TypedIndex(self._node_frames)

    @property
    def edges(self):
        return _TypedIndex(self._edge_frames, self.to_canonical_etype)

    def sharded_block_on(self, device, ranges: Dict[str, Tuple[int, int]], edge_weight: Optional[str] = None,
                         bounds=None) -> Block:
        """This rank's SHARD of the full-graph block (multi-GPU, ``distributed.py``): for every relation only the CSR
        rows of the destination range ``ranges[dst ntype] = (begin, end)`` -- ``indptr`` has ``end - begin + 1``
        entries and starts at 0, ``indices`` are GLOBAL source ids, ``eperm`` holds the global edge id of each slot.
        Only this rank's edges travel to the device, so the resident graph is ~1/world of the full CSR
        (``Block.shard_ranges`` marks the block; ``HeteroGraphConv`` then computes exactly those rows). ``bounds``:
        ``{ntype: world + 1 boundaries}`` of the node types whose ranges are NOT the equal-row ``shard_range`` cut
        (``distributed.work_bounds``), recorded so that ``sharded_forward`` all-gathers them as unequal chunks."""
        key = (str(device), edge_weight, tuple(sorted(ranges.items())))
        if key not in self._dev_blocks:
            from . import ops
            rels = {}
            for c, (s, d) in self._edges.items():
                if s.shape[0] > INT32_MAX or max(self._num[c[0]], self._num[c[2]]) > INT32_MAX:
                    raise OverflowError('int32 CSR cannot index relation %r' % (c,))
                b, e = ranges[c[2]]
                sel = np.nonzero((d >= b) & (d < e))[0]            # ascending edge ids: the stable order survives
                src = torch.from_numpy(s[sel].astype(np.int32, copy=False)).to(device)
                dst = torch.from_numpy((d[sel] - b).astype(np.int32, copy=False)).to(device)
                indptr, indices, eperm = ops.csr_build(src, dst, e - b)
                eid = torch.from_numpy(sel.astype(np.int64)).to(device)[eperm.long()]
                w = None
                if edge_weight is not None and edge_weight in self._edge_frames[c]:
                    w = self._edge_frames[c][edge_weight].reshape(-1)[eid.cpu()].to(device).to(torch.float32).contiguous()
                rels[c] = Relation(indptr, indices, self._num[c[0]], e - b, eid.to(torch.int32), w)
            blk = Block(rels, self._num, self._num)
            blk.shard_ranges = {t: (int(b), int(e)) for t, (b, e) in ranges.items()}
            blk.shard_bounds = {t: [int(x) for x in v] for t, v in (bounds or {}).items()}  # types cut by work (unequal)
            self._dev_blocks[key] = blk
        return self._dev_blocks[key]

    def device_edges(self, etype, device):
        u, v = self._edges[self.to_canonical_etype(etype)]
        if u.device != torch.device(device):
            u, v = u.to(device), v.to(device)
        return u, v

    def edge_arrays(self, etype):
        """Host copies (int64 numpy) -- for inspection and tests; the scoring kernel reads ``device_edges``."""
        u, v = self._edges[self.to_canonical_etype(etype)]
        return u.cpu().numpy().astype(np.int64), v.cpu().numpy().astype(np.int64)

    def all_edges(self, form='uv', order=None, etype=None):
        u, v = self._edges[self.to_canonical_etype(etype)]
        return u.long(), v.long()

    def to(self, device, **kwargs):
        return self
