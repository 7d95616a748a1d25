"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A minimal pure-PyTorch stand-in for the slice of ``dgl==0.5.2`` that the reference's hot
path touches, so that the reference's own ``src/model.py``,
``src/train/run.py::get_embeddings`` and ``src/metrics.py::get_recs`` can be imported
*verbatim* from ``/root/reference`` and executed on CPU (``tests/golden/make_golden.py``).

DGL 0.5.2 itself is not installable in this image (no wheel, no network), so every rule
below is a restatement of DGL's published semantics ("[DGL-recall]" in SURVEY.md 8c):

  (i)   fn.mean  = fp32 sum over in-edges / clamp(in_degree, 1); zero-degree rows = 0
  (ii)  fn.max   = elementwise max over in-edges; zero-degree rows = 0
  (iii) fn.u_mul_e broadcasts a per-edge scalar over the feature dimension
  (iv)  HeteroGraphConv: see dgl/nn/pytorch/__init__.py
  (v)   to_block: dst nodes = seeds in given order; src nodes = dst nodes first, then
        unseen edge sources in first-appearance order (etype order, then edge order)
  (vi)  heterograph(): ntypes sorted, canonical etypes sorted as tuples, edge id = list
        position, num_nodes = max id + 1 over all relations touching the type, ids int64
  (vii) apply_edges(u_dot_v) -> [E, 1]; edata[name] on a multi-etype graph -> dict keyed by
        canonical etype (only etypes holding the field); single-etype graph -> tensor

PARITY STATUS: "unpinned" for these DGL rules (the reference holds no test or golden vector
for them); everything that is plain torch/numpy in the reference (argsort/filter logic of
get_recs, max_margin_loss, NodeEmbedding, the relu/L2-norm epilogue of ConvLayer) runs as the
reference's own unmodified code on top of this shim.
"""
from contextlib import contextmanager

import numpy as np
import torch

NID = '_ID'
EID = '_ID'


def _t64(x):
    if isinstance(x, torch.Tensor):
        return x.to(torch.int64).reshape(-1)
    return torch.as_tensor(np.asarray(x), dtype=torch.int64).reshape(-1)


class _TypeView:
    """``g.nodes[ntype]`` / ``g.edges[etype]`` -> object with ``.data`` (a plain dict frame)."""

    def __init__(self, frame):
        self.data = frame


class _NodeSpace:
    def __init__(self, g, which):
        self._g, self._which = g, which

    def __getitem__(self, ntype):
        frames = self._g._src_frames if self._which == 'src' else self._g._dst_frames
        return _TypeView(frames[ntype])


class _EdgeSpace:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, etype):
        return _TypeView(self._g._edge_frames[self._g.to_canonical_etype(etype)])


class _MultiFrameView:
    """``g.ndata`` / ``block.srcdata`` / ``g.edata``: typed dict view with DGL's
    single-type shortcut (one type -> tensors directly, several -> dict keyed by type)."""

    def __init__(self, frames):
        self._frames = frames  # ordered dict: type -> frame dict

    def __getitem__(self, key):
        if len(self._frames) == 1:
            return next(iter(self._frames.values()))[key]
        return {t: f[key] for t, f in self._frames.items() if key in f}

    def __setitem__(self, key, val):
        if len(self._frames) == 1:
            next(iter(self._frames.values()))[key] = val
        else:
            assert isinstance(val, dict), 'multi-type graph needs a dict of tensors'
            for t, v in val.items():
                self._frames[t][key] = v

    def __contains__(self, key):
        return any(key in f for f in self._frames.values())

    def keys(self):
        ks = []
        for f in self._frames.values():
            for k in f:
                if k not in ks:
                    ks.append(k)
        return ks


class DGLHeteroGraph:
    def __init__(self, edges, num_src, num_dst, is_block=False, src_frames=None, dst_frames=None,
                 edge_frames=None):
        # edges: dict canonical etype -> (src int64, dst int64); sorted tuple order (rule vi)
        self._edges = {c: (_t64(e[0]), _t64(e[1])) for c, e in sorted(edges.items())}
        self.is_block = is_block
        self._num_src = dict(num_src)
        self._num_dst = dict(num_dst)
        self._src_frames = src_frames if src_frames is not None else {t: {} for t in sorted(num_src)}
        if is_block:
            self._dst_frames = dst_frames if dst_frames is not None else {t: {} for t in sorted(num_dst)}
        else:
            self._dst_frames = self._src_frames
        self._edge_frames = edge_frames if edge_frames is not None else {c: {} for c in self._edges}

    # ---- metagraph ----
    @property
    def ntypes(self):
        return sorted(self._num_src)

    @property
    def srctypes(self):
        return sorted(self._num_src)

    @property
    def dsttypes(self):
        return sorted(self._num_dst)

    @property
    def canonical_etypes(self):
        return list(self._edges.keys())

    @property
    def etypes(self):
        return [c[1] for c in self._edges]

    def to_canonical_etype(self, etype):
        if etype is None:
            assert len(self._edges) == 1
            return next(iter(self._edges))
        if isinstance(etype, tuple):
            return etype
        hits = [c for c in self._edges if c[1] == etype]
        if len(hits) != 1:
            raise KeyError(etype)
        return hits[0]

    # ---- sizes ----
    def num_nodes(self, ntype=None):
        if ntype is None:
            return sum(self._num_src.values())
        return self._num_src[ntype]

    number_of_nodes = num_nodes

    def number_of_src_nodes(self, ntype=None):
        return self.num_nodes(ntype)

    def number_of_dst_nodes(self, ntype=None):
        if ntype is None:
            return sum(self._num_dst.values())
        return self._num_dst[ntype]

    def num_edges(self, etype=None):
        if etype is None:
            return sum(int(e[0].numel()) for e in self._edges.values())
        return int(self._edges[self.to_canonical_etype(etype)][0].numel())

    number_of_edges = num_edges

    # ---- frames ----
    @property
    def nodes(self):
        return _NodeSpace(self, 'src')

    @property
    def srcnodes(self):
        return _NodeSpace(self, 'src')

    @property
    def dstnodes(self):
        return _NodeSpace(self, 'dst')

    @property
    def edges_view(self):
        return _EdgeSpace(self)

    @property
    def ndata(self):
        return _MultiFrameView(self._src_frames)

    @property
    def srcdata(self):
        return _MultiFrameView(self._src_frames)

    @property
    def dstdata(self):
        return _MultiFrameView(self._dst_frames)

    @property
    def edata(self):
        return _MultiFrameView(self._edge_frames)

    # ``g.edges[etype].data`` and ``g.edges(etype=...)`` are both used by the reference.
    class _EdgesAccessor:
        def __init__(self, g):
            self._g = g

        def __getitem__(self, etype):
            return _EdgeSpace(self._g)[etype]

        def __call__(self, etype=None, form='uv'):
            return self._g.all_edges(etype=etype, form=form)

    @property
    def edges(self):
        return DGLHeteroGraph._EdgesAccessor(self)

    # ---- structure queries ----
    def all_edges(self, form='uv', order=None, etype=None):
        s, d = self._edges[self.to_canonical_etype(etype)]
        if form == 'uv':
            return s, d
        if form == 'eid':
            return torch.arange(s.numel())
        return s, d, torch.arange(s.numel())

    def find_edges(self, eid, etype=None):
        s, d = self._edges[self.to_canonical_etype(etype)]
        eid = _t64(eid)
        return s[eid], d[eid]

    def out_edges(self, u, form='uv', etype=None):
        s, d = self._edges[self.to_canonical_etype(etype)]
        mask = torch.zeros(self._num_src[self.to_canonical_etype(etype)[0]], dtype=torch.bool)
        mask[_t64(u)] = True
        eid = torch.nonzero(mask[s]).reshape(-1)
        if form == 'eid':
            return eid
        if form == 'uv':
            return s[eid], d[eid]
        return s[eid], d[eid], eid

    def has_edges_between(self, u, v, etype=None):
        c = self.to_canonical_etype(etype)
        s, d = self._edges[c]
        key = set((s * self._num_dst.get(c[2], self._num_src[c[2]]) + d).tolist())
        q = (_t64(u) * self._num_dst.get(c[2], self._num_src[c[2]]) + _t64(v)).tolist()
        return torch.tensor([x in key for x in q], dtype=torch.bool)

    def in_degrees(self, etype=None):
        c = self.to_canonical_etype(etype)
        return torch.bincount(self._edges[c][1], minlength=self.number_of_dst_nodes(c[2]))

    # ---- relation slicing: g[stype, etype, dtype] ----
    def __getitem__(self, key):
        c = self.to_canonical_etype(key if isinstance(key, str) else tuple(key))
        s, d = self._edges[c]
        return DGLHeteroGraph({c: (s, d)},
                              {c[0]: self._num_src[c[0]]},
                              {c[2]: self.number_of_dst_nodes(c[2]) if self.is_block else self._num_src[c[2]]},
                              is_block=True,  # separate src/dst frames, like a sliced bipartite relation
                              edge_frames={c: self._edge_frames[c]})

    # ---- scope / device ----
    @contextmanager
    def local_scope(self):
        snap_s = {t: dict(f) for t, f in self._src_frames.items()}
        snap_d = None if self._dst_frames is self._src_frames else {t: dict(f) for t, f in self._dst_frames.items()}
        snap_e = {c: dict(f) for c, f in self._edge_frames.items()}
        try:
            yield
        finally:
            for t, f in self._src_frames.items():
                f.clear(); f.update(snap_s[t])
            if snap_d is not None:
                for t, f in self._dst_frames.items():
                    f.clear(); f.update(snap_d[t])
            for c, f in self._edge_frames.items():
                f.clear(); f.update(snap_e[c])

    def to(self, device, **kwargs):
        return self

    # ---- message passing (rules i-iii, vii) ----
    def update_all(self, message_func, reduce_func, etype=None):
        c = self.to_canonical_etype(etype)
        s, d = self._edges[c]
        n_dst = self.number_of_dst_nodes(c[2])
        x = self._src_frames[c[0]][message_func.lhs]
        if message_func.kind == 'copy_u':
            m = x[s]
        elif message_func.kind == 'u_mul_e':
            m = x[s] * self._edge_frames[c][message_func.rhs]
        else:
            raise NotImplementedError(message_func.kind)
        if not hasattr(reduce_func, 'kind'):
            raise NotImplementedError('UDF reducers (lstm) are outside the hot path (SURVEY.md 2.1 #1)')
        out = torch.zeros((n_dst,) + tuple(m.shape[1:]), dtype=m.dtype)
        if reduce_func.kind == 'mean':
            out.index_add_(0, d, m)
            deg = torch.bincount(d, minlength=n_dst).clamp(min=1).to(m.dtype)
            out = out / deg.reshape(-1, *([1] * (m.dim() - 1)))
        elif reduce_func.kind == 'sum':
            out.index_add_(0, d, m)
        elif reduce_func.kind == 'max':
            idx = d.reshape(-1, *([1] * (m.dim() - 1))).expand_as(m)
            out = out.scatter_reduce(0, idx, m, reduce='amax', include_self=False)
        else:
            raise NotImplementedError(reduce_func.kind)
        self._dst_frames[c[2]][reduce_func.out] = out

    def apply_edges(self, func, etype=None):
        c = self.to_canonical_etype(etype)
        s, d = self._edges[c]
        assert func.kind == 'u_dot_v'
        a = self._src_frames[c[0]][func.lhs]
        b = self._dst_frames[c[2]][func.rhs]
        self._edge_frames[c][func.out] = (a[s] * b[d]).sum(-1, keepdim=True)


def heterograph(data_dict, num_nodes_dict=None):
    """Rule (vi). ``data_dict``: canonical etype -> list of (src, dst) tuples or (src, dst) arrays."""
    edges = {}
    for c, data in data_dict.items():
        if isinstance(data, tuple) and len(data) == 2 and not isinstance(data[0], (int, np.integer)):
            s, d = data
        else:
            arr = np.asarray(list(data), dtype=np.int64).reshape(-1, 2)
            s, d = arr[:, 0], arr[:, 1]
        edges[c] = (_t64(s), _t64(d))
    num = {}
    for (st, _, dt), (s, d) in edges.items():
        num[st] = max(num.get(st, 0), int(s.max()) + 1 if s.numel() else 0)
        num[dt] = max(num.get(dt, 0), int(d.max()) + 1 if d.numel() else 0)
    if num_nodes_dict is not None:
        num.update(num_nodes_dict)
    return DGLHeteroGraph(edges, num, num, is_block=False)


def in_subgraph(g, nodes):
    """All in-edges of ``nodes`` (dict ntype -> ids); node space unchanged; keeps parent edge ids."""
    edges, eframes = {}, {}
    for c, (s, d) in g._edges.items():
        seeds = nodes.get(c[2])
        if seeds is None or len(seeds) == 0:
            keep = torch.zeros(0, dtype=torch.int64)
        else:
            mask = torch.zeros(g._num_src[c[2]], dtype=torch.bool)
            mask[_t64(seeds)] = True
            keep = torch.nonzero(mask[d]).reshape(-1)
        edges[c] = (s[keep], d[keep])
        eframes[c] = {k: v[keep] for k, v in g._edge_frames[c].items()}
        eframes[c][EID] = keep
    sg = DGLHeteroGraph(edges, g._num_src, g._num_src, is_block=False, src_frames=g._src_frames,
                        edge_frames=eframes)
    return sg


def to_block(g, dst_nodes):
    """Rule (v). ``g`` is a frontier on the parent's node space; features are row-gathers by NID."""
    dst_nodes = {t: _t64(v) for t, v in dst_nodes.items()}
    src_ids = {t: list(v.tolist()) for t, v in dst_nodes.items()}
    seen = {t: {int(n): i for i, n in enumerate(ids)} for t, ids in src_ids.items()}
    for c, (s, d) in g._edges.items():
        ids = src_ids.setdefault(c[0], [])
        sn = seen.setdefault(c[0], {})
        for u in s.tolist():
            if u not in sn:
                sn[u] = len(ids)
                ids.append(u)
    edges, eframes = {}, {}
    for c, (s, d) in g._edges.items():
        dn = {int(n): i for i, n in enumerate(dst_nodes.get(c[2], torch.zeros(0, dtype=torch.int64)).tolist())}
        ls = torch.tensor([seen[c[0]][int(u)] for u in s.tolist()], dtype=torch.int64)
        ld = torch.tensor([dn[int(v)] for v in d.tolist()], dtype=torch.int64)
        edges[c] = (ls, ld)
        eframes[c] = dict(g._edge_frames[c])
    num_src = {t: len(src_ids.get(t, [])) for t in g.ntypes}
    num_dst = {t: int(dst_nodes[t].numel()) if t in dst_nodes else 0 for t in g.ntypes}
    sf, df = {}, {}
    for t in g.ntypes:
        sid = torch.tensor(src_ids.get(t, []), dtype=torch.int64)
        did = dst_nodes.get(t, torch.zeros(0, dtype=torch.int64))
        sf[t] = {k: v[sid] for k, v in g._src_frames[t].items()}
        sf[t][NID] = sid
        df[t] = {k: v[did] for k, v in g._src_frames[t].items()}
        df[t][NID] = did
    return DGLHeteroGraph(edges, num_src, num_dst, is_block=True, src_frames=sf, dst_frames=df,
                          edge_frames=eframes)


from . import function  # noqa: E402,F401
from . import nn  # noqa: E402,F401
from . import dataloading  # noqa: E402,F401
