"""All-users x all-items cosine scoring + top-k on the device (the engine behind ``metrics.get_recs``).

Pipeline (every stage is a C-ABI call, see include/gnn_recsys_b200.h):

  colmean -> prep(items), prep(users) -> pass 1: tcgen05 GEMM + fused shortlist (ONE fp16 product, shortlist 32)
          -> exact fp32 re-score + soundness proof per user
          -> pass 2 for the users pass 1 could not prove: the 3-product hi/lo GEMM (fp32-grade) + re-score + proof
          -> brute-force fp32 kernel for whoever is left (exact ties filling a whole shortlist)

The GEMM runs on 16-bit operands; the answer does not: the final ids and their order always come from fp32
cosines (the torch formula the reference calls), and a user only keeps a shortlist answer when the MEASURED rounding
residuals of the operand rows prove that no item outside the shortlist can enter its top-k by more than ``tie_tol``
-- the tolerance below which the parity contract (and ``np.argsort`` in the reference) treats scores as tied.
On the BASELINE graphs pass 1 proves > 99 % of the users (tests/experiments/exp_two_product.py), so the tensor pipe
executes ~1.02 products per useful one instead of 3.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N
from . import ops

COS_EPS = 1e-6  # nn.CosineSimilarity(dim=1, eps=1e-6), src/metrics.py:58


@dataclass
class RecsConfig:
    elem: str = 'fp16'            # 16-bit operand type of the tensor-core GEMM: 'fp16' | 'bf16'
    parts_users: int = 1          # parts per user row: 1 = hi, 2 = hi + lo
    parts_items: int = 1          # parts per item row; (1,1) 1 product, (2,1) 2 products, (2,2) 3 products
    shortlist: int = 32           # candidates kept per user by the GEMM epilogue (>= k, <= 32)
    second: Optional[Tuple[str, int, int, int]] = ('fp16', 2, 2, 16)  # (elem, parts_users, parts_items, shortlist) of
    #                               pass 2 over the users pass 1 cannot prove; None = straight to the exact kernel
    center: bool = True           # subtract the mean normalised item row (ranking-invariant, shrinks the error bound)
    k_band: bool = True           # epilogue threshold max(S-th best, k-th best - 2 err) instead of the S-th best alone
    tie_tol: float = 1e-5         # score gap treated as a tie (north_star parity rule); 0 = strict
    acc_err: float = 1.5e-6       # allowance for fp32 accumulation error of the tensor-core sum, relative to |x||y|
    exact_only: bool = False      # skip the tensor-core path (brute-force fp32 kernel for every user)
    single_cta: bool = False      # cta_group::1 kernel instead of CTA pairs (same results; tests / experiments)
    tail_split: bool = True       # cut the user tiles of the last, partial wave into item ranges (wave quantisation)
    small_items: int = 32768      # calls with fewer items than this AND fewer than small_work scores take the 3-product
    small_work: int = 1 << 31     # scheme directly with its 16-entry shortlist: a sweep that short is all shortlist
    #                               warm-up, and the single product's proof would only add a second pass (and a host
    #                               read) to a launch-bound step (c1: 10k x 5k)
    item_order: bool = True       # sweep the items in descending order of their cosine to the mean user row (thresholds
    #                               settle after the first tiles instead of climbing through the whole sweep; same result)
    order_min_items: int = 16384  # ... for tables at least this long and calls of at least order_min_work scores
    order_min_work: int = 1 << 31
    parts: Optional[int] = None   # shorthand: parts=1 -> (1, 1), parts=2 -> (2, 2) and no second pass

    def __post_init__(self):
        if self.parts is not None:
            self.parts_users = self.parts_items = int(self.parts)
            if self.parts == 2:
                self.second = None
        if (self.parts_users, self.parts_items) not in ((1, 1), (2, 1), (2, 2)):
            raise ValueError('(parts_users, parts_items) must be (1, 1), (2, 1) or (2, 2)')
        if self.second is not None and (self.second[1], self.second[2]) == (self.parts_users, self.parts_items) \
                and self.second[0] == self.elem:
            self.second = None

    @property
    def elem_type(self) -> int:
        return elem_type_of(self.elem)

    @property
    def products(self) -> int:
        return {(1, 1): 1, (2, 1): 2, (2, 2): 3}[(self.parts_users, self.parts_items)]

    @property
    def flags(self) -> int:
        return (N.SCORE_FLAG_SINGLE_CTA if self.single_cta else 0) | (0 if self.tail_split else N.SCORE_FLAG_NO_TAIL_SPLIT)


def elem_type_of(elem: str) -> int:
    return {'bf16': N.ELEM_BF16, 'fp16': N.ELEM_FP16}[elem]


def score_err_bound(ru: float, ru1: float, item_stats, elem: str, parts_users: int, parts_items: int,
                    acc_err: float) -> float:
    """Host mirror of the device formula (``score_err`` in csrc/topk_aux.cu): |approximate - exact| of any score of a
    user whose operand row has rounding residuals ``ru`` (final) / ``ru1`` (first level) against an item table with
    statistics ``item_stats = [Y, min |x|, R, R1]``:
    ``s - a = (x - xq).y + xq.(y - yq) [+ x_lo.y_lo for the 3-product scheme] + accumulation``."""
    Y, _, R, R1 = [float(v) for v in item_stats]
    u = 2.0 ** -11 if elem == 'fp16' else 2.0 ** -8
    err = ru * Y + (1 + ru) * R + acc_err * (1 + ru) * (Y + R)
    if parts_users == 2 and parts_items == 2:
        err += ru1 * R1 * (1 + u) ** 2
    return err


class BoughtCSR:
    """Already-bought items per user as a CSR (int64 indptr, int32 ids sorted ascending within a row) -- the
    device-side form of the reference's ``already_bought_dict`` (``src/metrics.py:19-28``). Indexing with a
    user id returns that user's list, so it can stand in for the dict at the reference's call sites."""

    def __init__(self, indptr: np.ndarray, ids: np.ndarray, row_of_user: Optional[dict] = None):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._row_of_user = row_of_user  # None: row == user id
        self._dev = {}

    @property
    def n_rows(self) -> int:
        return self.indptr.shape[0] - 1

    @classmethod
    def from_edges(cls, users, items, n_users: int) -> 'BoughtCSR':
        users = np.asarray(users).astype(np.int64).reshape(-1)
        items = np.asarray(items).astype(np.int64).reshape(-1)
        order = np.lexsort((items, users))
        indptr = np.zeros(n_users + 1, dtype=np.int64)
        if users.size:
            np.cumsum(np.bincount(users, minlength=n_users), out=indptr[1:])
        return cls(indptr, items[order].astype(np.int32))

    @classmethod
    def from_dict(cls, d, user_ids) -> 'BoughtCSR':
        """Rows follow ``user_ids`` order; users missing from a plain dict have no purchases."""
        has_default = hasattr(d, '__missing__')
        lists = []
        for u in user_ids:
            lst = d[u] if (has_default or u in d) else []
            lists.append(np.sort(np.asarray(lst, dtype=np.int64).reshape(-1)))
        lens = np.fromiter((a.size for a in lists), dtype=np.int64, count=len(lists))
        indptr = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum(lens, out=indptr[1:])
        ids = np.concatenate(lists) if lists and indptr[-1] > 0 else np.zeros(0, dtype=np.int64)
        return cls(indptr, ids.astype(np.int32), {int(u): r for r, u in enumerate(user_ids)})

    def select(self, user_ids) -> 'BoughtCSR':
        """Rows re-ordered to follow ``user_ids`` (vectorised)."""
        if isinstance(user_ids, range) and self._row_of_user is None and user_ids.step == 1:
            lo, hi = user_ids.start, user_ids.stop  # contiguous slice: no gather
            return BoughtCSR(self.indptr[lo:hi + 1] - self.indptr[lo], self.ids[self.indptr[lo]:self.indptr[hi]])
        uid = np.asarray(user_ids, dtype=np.int64).reshape(-1)
        if self._row_of_user is not None:
            uid = np.asarray([self._row_of_user[int(u)] for u in uid], dtype=np.int64)
        if uid.size == self.n_rows and np.array_equal(uid, np.arange(self.n_rows)):
            return self
        lo, hi = self.indptr[uid], self.indptr[uid + 1]
        lens = hi - lo
        indptr = np.zeros(uid.size + 1, dtype=np.int64)
        np.cumsum(lens, out=indptr[1:])
        take = np.repeat(lo - indptr[:-1], lens) + np.arange(int(indptr[-1]))
        return BoughtCSR(indptr, self.ids[take])

    def __getitem__(self, user):
        r = int(user) if self._row_of_user is None else self._row_of_user[int(user)]
        return self.ids[self.indptr[r]:self.indptr[r + 1]].tolist()

    def on(self, device):
        key = str(device)
        if key not in self._dev:
            ids = self.ids if self.ids.size else np.zeros(1, dtype=np.int32)  # keep a valid pointer
            self._dev[key] = (torch.from_numpy(self.indptr).to(device), torch.from_numpy(ids).to(device))
        return self._dev[key]


class ScoringTable:
    """Item side of the scoring GEMM, prepared once per embedding table: quantised rows, centre, residual stats.
    ``operands(elem, parts)`` prepares (and caches) the item operand of another scheme on demand (pass 2)."""

    def __init__(self, h_item: torch.Tensor, cfg: RecsConfig, item_id_base: int = 0):
        self.h_item = h_item.contiguous()
        self.cfg, self.item_id_base = cfg, int(item_id_base)
        self.n_items, self.d = h_item.shape
        self.tc = (not cfg.exact_only) and self.d <= 256 and self.n_items > 0
        self.center = None
        self._ops = {}
        if self.tc:
            self.d_pad = 64 * ((self.d + 63) // 64)
            if self.d_pad > 128:  # out_dim 192 / 256 presets (main.py:86): single product only, no tensor-core second pass
                from dataclasses import replace
                self.cfg = cfg = replace(cfg, parts=None, parts_users=1, parts_items=1, second=None)
            if cfg.center:
                self.center = ops.colmean_normalized(self.h_item)

    @property
    def items_q(self):  # operand rows / residual statistics of the configuration's own (first-pass) scheme
        return self.operands(self.cfg.elem, self.cfg.parts_items)[0] if self.tc else None

    @property
    def stats(self):
        return self.operands(self.cfg.elem, self.cfg.parts_items)[1] if self.tc else None

    def operands(self, elem: str, parts: int):
        key = (elem, parts)
        if key not in self._ops:
            self._ops[key] = ops.score_prep(self.h_item, self.center, self.d_pad, parts, elem_type_of(elem), True)
        return self._ops[key]


def _tc_pass(h_user, table: ScoringTable, k: int, bptr, bids, scheme, cfg: RecsConfig, user_map=None, mark=None,
             order=None):
    """One tensor-core pass: prep(users) -> GEMM + shortlist -> exact re-score + proof.
    Returns ``(ids, scores, overflow list, n_overflow)``; overflow entries are ``user_map`` values when given.
    ``order``: item sweep order (``ops.score_item_order``) or None for item-id order."""
    elem, pu, pi, shortlist = scheme
    et = elem_type_of(elem)
    shortlist = max(shortlist, k)
    items_q, item_stats = table.operands(elem, pi)
    if order is not None:
        items_q = ops.permute_rows(items_q, order)
    users_q, user_stats = ops.score_prep(h_user, None, table.d_pad, pu, et, cfg.k_band)
    band = ops.score_band(item_stats, user_stats, et, pu, pi, cfg.acc_err) if cfg.k_band else None
    if mark:
        mark('score_begin')
    sl_score, sl_id = ops.score_topk_tc(users_q, items_q, table.item_id_base, table.d_pad, pu, pi, et, bptr, bids,
                                        shortlist, k, band, user_map, cfg.flags, item_perm=order)
    if mark:
        mark('score_end')
    return ops.rescore_topk(h_user, table.h_item, table.item_id_base, table.center, sl_score, sl_id, item_stats, et, pu,
                            pi, cfg.acc_err, band, cfg.tie_tol, k, COS_EPS, user_map)


def recommend_topk(h_user: torch.Tensor, table: ScoringTable, k: int, bought: Optional[BoughtCSR] = None,
                   return_overflow: bool = False, mark=None, popularity: Optional[torch.Tensor] = None,
                   weight_popularity: float = 1.0):
    """Top-``k`` items of ``table`` for every row of ``h_user`` (``[n, d]`` fp32, CUDA): ``(ids int32 [n, k],
    scores fp32 [n, k])`` sorted by (score desc, id asc); ``-1`` / ``-inf`` pad rows with fewer than k candidates.
    ``bought`` rows must follow ``h_user`` rows. ids are global (``table.item_id_base`` added).
    ``mark(name)`` (optional) is called between stages -- bench.py records CUDA events with it.
    ``return_overflow``: also return ``(users sent to pass 2, users sent to the exact kernel)`` as Python ints."""
    mark = mark or (lambda name: None)
    cfg = table.cfg
    h_user = h_user.contiguous()
    n = h_user.shape[0]
    dev = h_user.device
    bptr, bids = (None, None) if bought is None else bought.on(dev)
    if bought is not None and bought.n_rows != n:
        raise ValueError('bought rows (%d) do not match user rows (%d)' % (bought.n_rows, n))
    if popularity is not None or not table.tc or k > 32:
        # popularity re-rank (softmax over ALL items + w * popularity), embeddings wider than 128 and k > 32 (the fused
        # epilogue keeps at most 32 candidates) run on the exact fp32 kernel: any k, one item sweep per 64 entries
        ids, scores = ops.score_topk_exact(h_user, table.h_item, table.item_id_base, bptr, bids, k, COS_EPS,
                                           popularity=popularity, weight_popularity=weight_popularity)
        return (ids, scores, (0, 0)) if return_overflow else (ids, scores)
    if cfg.shortlist > 32:
        raise ValueError('shortlist above 32 is not supported by the fused top-k epilogue')
    first, second = (cfg.elem, cfg.parts_users, cfg.parts_items, cfg.shortlist), cfg.second
    if second is not None and cfg.products < 3 and table.n_items < cfg.small_items and n * table.n_items < cfg.small_work:
        first, second = second, None   # short sweep: the fp32-grade scheme directly (see RecsConfig.small_items)
    order = None
    if cfg.item_order and n > 0 and table.n_items >= max(cfg.order_min_items, 1) and n * table.n_items >= cfg.order_min_work:
        # sweep order of this call: items by descending cosine to the mean normalised user row (shared by both passes)
        order = ops.score_item_order(table.h_item, ops.colmean_normalized(h_user))
    ids, scores, overflow, n_overflow = _tc_pass(h_user, table, k, bptr, bids, first, cfg, mark=mark, order=order)
    mark('rescore_end')
    n1 = n2 = 0
    if second is None:
        # no tensor-core second pass: the exact kernel reads the overflow count on the device (no host read at all)
        if n > 0:
            ops.score_topk_exact(h_user, table.h_item, table.item_id_base, bptr, bids, k, COS_EPS, user_list=overflow,
                                 n_list=n_overflow, out_ids=ids, out_scores=scores)
            if return_overflow:
                n1 = n2 = int(n_overflow.item())
    else:
        if n > 0:
            n1 = int(n_overflow.item())  # the one host read of the pipeline: sizes pass 2 (usually < 1 % of the users)
        if n1 > 0:
            # pass 2: the users pass 1 could not prove, compacted, through the more accurate scheme
            rows = overflow[:n1].sort().values  # ascending: deterministic whatever order the proof kernel appended in
            # (the sweep order pays for one permutation lookup per candidate: only worth it on a long pass)
            ids2, scores2, overflow, n_overflow = _tc_pass(h_user[rows.long()], table, k, bptr, bids, second, cfg,
                                                           user_map=rows,
                                                           order=order if n1 * table.n_items >= cfg.order_min_work else None)
            ids[rows.long()] = ids2
            scores[rows.long()] = scores2
            # exact fp32 pass for whoever is left (device-side list and count)
            ops.score_topk_exact(h_user, table.h_item, table.item_id_base, bptr, bids, k, COS_EPS, user_list=overflow,
                                 n_list=n_overflow, out_ids=ids, out_scores=scores)
            if return_overflow:
                n2 = int(n_overflow.item())
    mark('fallback_end')
    return (ids, scores, (n1, n2)) if return_overflow else (ids, scores)
