"""Thin Python wrappers over the C ABI: one function per entry point, tensors in / tensors out.

Everything here runs on the current CUDA device and stream; inputs must be contiguous CUDA tensors (fp32
features, int32 CSR). No CPU path exists -- see ``_native.py``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native as N

_ws_cache = {}


def _ws(nbytes: int, device, tag: str) -> torch.Tensor:
    """Per-(device, tag) grow-only scratch buffer (the library never allocates)."""
    key = (str(device), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = N.workspace(nbytes, device)
        _ws_cache[key] = buf
    return buf


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def linear(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
           legacy: bool = False) -> torch.Tensor:
    """``y = x @ wt (+ bias) (relu)`` with ``wt`` = ``nn.Linear.weight.t()`` ([d_in, d_out], contiguous).
    ``legacy=True`` keeps the square bias-free shapes off the tcgen05 path (3xTF32 mma.sync instead; tests / experiments)."""
    x, wt = _f32(x), _f32(wt)
    n, d_in = x.shape
    d_out = wt.shape[1]
    assert wt.shape[0] == d_in
    y = torch.empty((n, d_out), dtype=torch.float32, device=x.device)
    ws = _ws(N.load().gr_linear_workspace_bytes(n, d_in, d_out), x.device, 'linear')
    N.call('gr_linear_f32', N.ptr(x), n, d_in, N.ptr(wt), N.ptr(_f32(bias)) if bias is not None else None, d_out,
           int(relu), N.LINEAR_FLAG_LEGACY if legacy else 0, N.ptr(y), N.ptr(ws), ws.numel(), N.stream())
    return y


def sage_relation(indptr, indices, edge_w, h_src, h_dst, w_self_t, w_neigh_t, out, reducer: int, l2norm: bool,
                  accumulate: int = N.ACC_STORE, z_scale: float = 1.0, row_begin: int = 0, row_end: Optional[int] = None,
                  flags: int = 0, packed: Optional[torch.Tensor] = None):
    """Fused ``ConvLayer.forward`` of one relation into ``out[row_begin:row_end]`` (see include/gnn_recsys_b200.h)."""
    nnz = int(indices.shape[0])
    n_dst = int(indptr.shape[0]) - 1
    row_end = n_dst if row_end is None else row_end
    d_neigh, d_self, d_out = h_src.shape[1], h_dst.shape[1], w_self_t.shape[1]
    assert w_self_t.shape[0] == d_self and w_neigh_t.shape[0] == d_neigh and w_neigh_t.shape[1] == d_out
    assert out.shape[1] == d_out and indptr.dtype == torch.int32 and indices.dtype == torch.int32
    lib = N.load()
    nb = lib.gr_sage_relation_workspace_bytes(nnz, d_neigh)
    ws = _ws(nb, out.device, 'sage')
    N.call('gr_sage_relation_f32', N.ptr(indptr), N.ptr(indices), N.ptr(edge_w) if edge_w is not None else None, nnz,
           N.ptr(h_src), N.ptr(h_dst), row_begin, row_end, d_neigh, d_self, N.ptr(w_self_t), N.ptr(w_neigh_t), d_out,
           reducer, int(l2norm), accumulate, float(z_scale), int(flags), N.ptr(packed) if packed is not None else None,
           N.ptr(out), N.ptr(ws), ws.numel(), N.stream())
    return out


def sage_pack_weights(w_self_t: torch.Tensor, w_neigh_t: torch.Tensor, flags: int = 0) -> Optional[torch.Tensor]:
    """The two projection matrices in the form the fused ConvLayer kernel reads (``gr_sage_pack_weights``); ``None`` when
    the dimensions take the generic kernel. Pack once per weight update and pass as ``sage_relation(..., packed=)``."""
    d_self, d_out = w_self_t.shape
    d_neigh = w_neigh_t.shape[0]
    nb = N.load().gr_sage_packed_weights_bytes(d_neigh, d_self, d_out)
    if nb == 0:
        return None
    packed = N.workspace(nb, w_self_t.device)
    N.call('gr_sage_pack_weights', N.ptr(w_self_t), N.ptr(w_neigh_t), d_neigh, d_self, d_out, int(flags), N.ptr(packed),
           N.stream())
    return packed


def gather_reduce(indptr, indices, edge_w, h_src, reducer: int, row_begin: int = 0, row_end: Optional[int] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    nnz = int(indices.shape[0])
    n_dst = int(indptr.shape[0]) - 1
    row_end = n_dst if row_end is None else row_end
    d = h_src.shape[1]
    if out is None:
        out = torch.zeros((n_dst, d), dtype=torch.float32, device=h_src.device)
    lib = N.load()
    ws = _ws(lib.gr_sage_relation_workspace_bytes(nnz, d), h_src.device, 'sage')
    N.call('gr_gather_reduce_f32', N.ptr(indptr), N.ptr(indices), N.ptr(edge_w) if edge_w is not None else None, nnz,
           N.ptr(h_src), row_begin, row_end, d, reducer, N.ptr(out), N.ptr(ws), ws.numel(), N.stream())
    return out


def edge_cosine(u: torch.Tensor, v: torch.Tensor, h_src: torch.Tensor, h_dst: torch.Tensor) -> torch.Tensor:
    """Cosine of ``h_src[u[e]]`` and ``h_dst[v[e]]`` for every edge; returns ``[E, 1]`` like ``edata['cos']``."""
    e = int(u.shape[0])
    out = torch.empty((e, 1), dtype=torch.float32, device=h_src.device)
    assert u.dtype == torch.int32 and v.dtype == torch.int32 and h_src.shape[1] == h_dst.shape[1]
    N.call('gr_edge_cosine_f32', N.ptr(u), N.ptr(v), e, N.ptr(h_src), N.ptr(h_dst), h_src.shape[1], N.ptr(out),
           N.stream())
    return out


def colmean_normalized(x: torch.Tensor) -> torch.Tensor:
    n, d = x.shape
    center = torch.empty(d, dtype=torch.float32, device=x.device)
    lib = N.load()
    ws = _ws(lib.gr_colmean_workspace_bytes(n, d), x.device, 'colmean')
    N.call('gr_colmean_normalized_f32', N.ptr(x), n, d, N.ptr(center), N.ptr(ws), ws.numel(), N.stream())
    return center


def score_prep(x: torch.Tensor, center: Optional[torch.Tensor], d_pad: int, parts: int, elem_type: int,
               want_stats: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Quantised operand rows ``[n, parts * d_pad]`` (int16 storage) and ``stats = [max |row - center|, min |row|,
    max final rounding residual, max first-level residual]`` (see include/gnn_recsys_b200.h)."""
    n, d = x.shape
    q = torch.empty((n, parts * d_pad), dtype=torch.int16, device=x.device)
    stats = None
    if want_stats:  # {0, +inf, 0, 0} built on the device (no pageable H2D copy: capture-safe)
        stats = torch.zeros(4, dtype=torch.float32, device=x.device)
        stats[1] = float('inf')
    N.call('gr_score_prep', N.ptr(x), n, d, N.ptr(center) if center is not None else None, d_pad, parts, elem_type,
           N.ptr(q), N.ptr(stats) if stats is not None else None, N.stream())
    return q, stats


def score_topk_tc(users_q, items_q, item_id_base: int, d_pad: int, parts_users: int, parts_items: int, elem_type: int,
                  bought_indptr, bought_ids, shortlist: int, k: Optional[int] = None, band: Optional[torch.Tensor] = None,
                  user_map: Optional[torch.Tensor] = None, flags: int = 0, item_perm: Optional[torch.Tensor] = None):
    """``item_perm`` (int32 [n_items]): row p of ``items_q`` is item ``item_perm[p]`` (see ``score_item_order``)."""
    n_users, n_items = users_q.shape[0], items_q.shape[0]
    assert item_perm is None or (item_perm.dtype == torch.int32 and item_perm.numel() == n_items)
    dev = users_q.device
    sl_score = torch.empty((n_users, shortlist), dtype=torch.float32, device=dev)
    sl_id = torch.empty((n_users, shortlist), dtype=torch.int32, device=dev)
    lib = N.load()
    ws = _ws(lib.gr_score_topk_workspace_bytes(n_users, n_items, shortlist), dev, 'score')
    N.call('gr_score_topk_tc', N.ptr(users_q), n_users, N.ptr(items_q), n_items, item_id_base,
           N.ptr(item_perm) if item_perm is not None else None, d_pad, parts_users, parts_items, elem_type, N.ptr(bought_indptr) if bought_indptr is not None else None,
           N.ptr(bought_ids) if bought_ids is not None else None, shortlist, shortlist if k is None else k,
           N.ptr(band) if band is not None else None, N.ptr(user_map) if user_map is not None else None, flags,
           N.ptr(sl_score), N.ptr(sl_id), N.ptr(ws), ws.numel(), N.stream())
    return sl_score, sl_id


def score_item_order(h_item: torch.Tensor, direction: torch.Tensor) -> torch.Tensor:
    """int32 [n_items]: item indices in descending order of cos(h_item[i], direction) -- the sweep order that lets the
    shortlist thresholds of ``score_topk_tc`` settle after the first tiles (csrc/item_order.cu)."""
    n, d = h_item.shape
    assert h_item.dtype == torch.float32 and h_item.is_contiguous() and direction.numel() == d
    perm = torch.empty(n, dtype=torch.int32, device=h_item.device)
    lib = N.load()
    ws = _ws(lib.gr_score_item_order_workspace_bytes(n), h_item.device, 'order')
    N.call('gr_score_item_order', N.ptr(h_item), n, d, N.ptr(direction), N.ptr(perm), N.ptr(ws), ws.numel(), N.stream())
    return perm


def permute_rows(x: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """``out[p] = x[perm[p]]`` for a 2-D contiguous tensor whose rows are a multiple of 16 bytes."""
    assert x.dim() == 2 and x.is_contiguous() and perm.dtype == torch.int32 and perm.numel() == x.shape[0]
    out = torch.empty_like(x)
    N.call('gr_permute_rows', N.ptr(x), x.shape[0], x.shape[1] * x.element_size(), N.ptr(perm), N.ptr(out), N.stream())
    return out


def score_band(item_stats, user_stats, elem_type: int, parts_users: int, parts_items: int, acc_err: float) -> torch.Tensor:
    """Device scalar: 2 x the largest scoring error over the user table (the k-th-best margin of the epilogue)."""
    band = torch.empty(1, dtype=torch.float32, device=item_stats.device)
    N.call('gr_score_band', N.ptr(item_stats), N.ptr(user_stats), elem_type, parts_users, parts_items, float(acc_err),
           N.ptr(band), N.stream())
    return band


def rescore_topk(h_user, h_item, item_id_base: int, center, sl_score, sl_id, stats, elem_type: int, parts_users: int,
                 parts_items: int, acc_err: float, band: Optional[torch.Tensor], tie_tol: float, k: int, eps: float,
                 user_map: Optional[torch.Tensor] = None, overflow=None, n_overflow=None):
    """Exact re-score + proof. ``overflow`` / ``n_overflow`` may be passed in to APPEND to an existing list."""
    n_users, d = h_user.shape
    dev = h_user.device
    out_ids = torch.empty((n_users, k), dtype=torch.int32, device=dev)
    out_scores = torch.empty((n_users, k), dtype=torch.float32, device=dev)
    if overflow is None:
        overflow = torch.empty(max(n_users, 1), dtype=torch.int32, device=dev)
        n_overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    N.call('gr_rescore_topk_f32', N.ptr(h_user), N.ptr(h_item), item_id_base, d,
           N.ptr(center) if center is not None else None, N.ptr(sl_score), N.ptr(sl_id), sl_id.shape[1], n_users,
           N.ptr(stats), elem_type, parts_users, parts_items, float(acc_err), N.ptr(band) if band is not None else None,
           float(tie_tol), k, float(eps), N.ptr(user_map) if user_map is not None else None, N.ptr(out_ids),
           N.ptr(out_scores), N.ptr(overflow), N.ptr(n_overflow), N.stream())
    return out_ids, out_scores, overflow, n_overflow


def score_topk_exact(h_user, h_item, item_id_base: int, bought_indptr, bought_ids, k: int, eps: float,
                     user_list=None, n_list=None, out_ids=None, out_scores=None, popularity=None,
                     weight_popularity: float = 1.0):
    """Exact fp32 top-k for the listed users (all users when ``user_list`` is None), written in place into
    ``out_ids`` / ``out_scores`` rows of those users."""
    n_users, d = h_user.shape
    dev = h_user.device
    if out_ids is None:
        out_ids = torch.full((n_users, k), -1, dtype=torch.int32, device=dev)
        out_scores = torch.full((n_users, k), float('-inf'), dtype=torch.float32, device=dev)
    N.call('gr_score_topk_exact_f32', N.ptr(h_user), N.ptr(user_list) if user_list is not None else None,
           N.ptr(n_list) if n_list is not None else None, n_users, N.ptr(h_item), h_item.shape[0], item_id_base, d,
           N.ptr(bought_indptr) if bought_indptr is not None else None,
           N.ptr(bought_ids) if bought_ids is not None else None, k, eps,
           N.ptr(popularity) if popularity is not None else None, float(weight_popularity), N.ptr(out_ids),
           N.ptr(out_scores), N.stream())
    return out_ids, out_scores


def metrics_at_k(recs: torch.Tensor, truth_indptr: torch.Tensor, truth_ids: torch.Tensor, n_items: int) -> torch.Tensor:
    """Counters of ``recs_to_metrics`` (see include/gnn_recsys_b200.h): uint64[5] on the device (as int64)."""
    n_users, k = recs.shape
    dev = recs.device
    counters = torch.zeros(5, dtype=torch.int64, device=dev)
    ws = _ws(N.load().gr_metrics_workspace_bytes(n_items), dev, 'metrics')
    N.call('gr_metrics_at_k', N.ptr(recs), n_users, k, N.ptr(truth_indptr), N.ptr(truth_ids), n_items, N.ptr(counters),
           N.ptr(ws), ws.numel(), N.stream())
    return counters


def topk_merge(scores: torch.Tensor, ids: torch.Tensor, k_out: int):
    """``scores`` / ``ids``: ``[parts, n_users, k_in]`` -> ``[n_users, k_out]`` best by (score desc, id asc)."""
    parts, n_users, k_in = scores.shape
    dev = scores.device
    out_s = torch.empty((n_users, k_out), dtype=torch.float32, device=dev)
    out_i = torch.empty((n_users, k_out), dtype=torch.int32, device=dev)
    N.call('gr_topk_merge', N.ptr(scores), N.ptr(ids), parts, n_users, k_in, k_out, N.ptr(out_s), N.ptr(out_i),
           N.stream())
    return out_s, out_i


def csr_build(src: torch.Tensor, dst: torch.Tensor, n_dst: int, validate: bool = True):
    """Stable COO -> CSR over destination rows on the device: ``(indptr, indices, eperm)``, all int32.
    ``validate``: read back the device status word and raise ``IndexError`` for destination ids outside
    ``[0, n_dst)`` like the host twin ``graph.csr_by_dst_host`` (one 4-byte D2H; ingest is not a hot path)."""
    assert src.dtype == torch.int32 and dst.dtype == torch.int32
    nnz = int(src.shape[0])
    dev = src.device
    indptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    eperm = torch.empty(nnz, dtype=torch.int32, device=dev)
    lib = N.load()
    ws = N.workspace(lib.gr_csr_build_workspace_bytes(nnz, n_dst), dev)
    status = torch.empty(1, dtype=torch.int32, device=dev) if validate else None
    N.call('gr_csr_build_i32', N.ptr(src), N.ptr(dst), nnz, n_dst, N.ptr(indptr), N.ptr(indices), N.ptr(eperm),
           N.ptr(status) if validate else None, N.ptr(ws), ws.numel(), N.stream())
    if validate and int(status.item()) != 0:
        raise IndexError('destination ids outside [0, %d)' % n_dst)
    return indptr, indices, eperm


def remap_first_appearance(raw: torch.Tensor, defer: bool = False):
    """Raw int64 ids -> ``(new_ids int32 [n], uniq_raw int64 [n_unique])``: contiguous ids in order of first
    appearance (``create_ids``, reference ``src/builder.py:182-227``). ``defer=True`` skips the host read of the count
    and returns ``(new_ids, uniq_raw buffer [n], n_unique device int32 [1])``."""
    assert raw.dtype == torch.int64
    n = int(raw.shape[0])
    dev = raw.device
    new_ids = torch.empty(n, dtype=torch.int32, device=dev)
    uniq = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    n_unique = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = N.load()
    ws = N.workspace(lib.gr_remap_workspace_bytes(n), dev)
    N.call('gr_remap_first_appearance_i64', N.ptr(raw), n, N.ptr(new_ids), N.ptr(uniq), N.ptr(n_unique), N.ptr(ws),
           ws.numel(), N.stream())
    if defer:
        return new_ids, uniq, n_unique
    return new_ids, uniq[:int(n_unique.item())]


def remap_many(raws):
    """``remap_first_appearance`` of several id arrays with ONE host read of their unique counts."""
    outs = [remap_first_appearance(r, defer=True) for r in raws]
    if not outs:
        return []
    counts = torch.cat([o[2] for o in outs]).tolist()
    return [(o[0], o[1][:c]) for o, c in zip(outs, counts)]


def sample_count(indptr: torch.Tensor, eperm: Optional[torch.Tensor], seeds: torch.Tensor, fanout: int,
                 excl_sorted: Optional[torch.Tensor] = None):
    """Per-seed frontier sizes of one relation as the block's ``indptr`` (int32 ``[n_seeds + 1]``) plus the device
    scalar ``total`` (see include/gnn_recsys_b200.h, ``gr_sample_count_i32``)."""
    assert seeds.dtype == torch.int64 and indptr.dtype == torch.int32
    n = int(seeds.shape[0])
    dev = indptr.device
    out_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    total = torch.empty(1, dtype=torch.int32, device=dev)
    n_excl = 0 if excl_sorted is None else int(excl_sorted.shape[0])
    N.call('gr_sample_count_i32', N.ptr(indptr), N.ptr(eperm) if eperm is not None else None, N.ptr(seeds), n,
           int(fanout), N.ptr(excl_sorted) if n_excl else None, n_excl, N.ptr(out_indptr), N.ptr(total), N.stream())
    return out_indptr, total


def sample_fill(indptr, indices, eperm, seeds, fanout: int, excl_sorted, key: int, out_indptr, out_src: torch.Tensor,
                out_eid: torch.Tensor):
    """Writes the frontier's global source ids (int64) and edge ids (int32) into ``out_src`` / ``out_eid`` (views are
    fine: only their first ``total`` entries are written)."""
    assert out_src.dtype == torch.int64 and out_eid.dtype == torch.int32
    n_excl = 0 if excl_sorted is None else int(excl_sorted.shape[0])
    N.call('gr_sample_fill_i32', N.ptr(indptr), N.ptr(indices), N.ptr(eperm) if eperm is not None else None,
           N.ptr(seeds), int(seeds.shape[0]), int(fanout), N.ptr(excl_sorted) if n_excl else None, n_excl,
           int(key) & 0xFFFFFFFFFFFFFFFF, N.ptr(out_indptr), N.ptr(out_src), N.ptr(out_eid), N.stream())


def negative_uniform(edge_src: torch.Tensor, eids: torch.Tensor, k: int, n_dst_nodes: int, key: int):
    """``negative_sampler.Uniform(k)`` on the device: ``(src, dst)`` int64 ``[n_pos * k]``, k consecutive per edge."""
    assert edge_src.dtype == torch.int32 and eids.dtype == torch.int64
    n = int(eids.shape[0])
    dev = edge_src.device
    src = torch.empty(n * k, dtype=torch.int64, device=dev)
    dst = torch.empty(n * k, dtype=torch.int64, device=dev)
    N.call('gr_negative_uniform_i64', N.ptr(edge_src), N.ptr(eids), n, int(k), int(n_dst_nodes),
           int(key) & 0xFFFFFFFFFFFFFFFF, N.ptr(src), N.ptr(dst), N.stream())
    return src, dst
