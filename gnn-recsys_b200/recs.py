"""All-users x all-items cosine scoring + top-k on the device (the engine behind ``metrics.get_recs``).

Pipeline (every stage is a C-ABI call, see include/gnn_recsys_b200.h):

  colmean -> prep(items), prep(users) -> tcgen05 GEMM + fused shortlist -> exact fp32 re-score + soundness proof
          -> exact fallback for the (rare) users whose shortlist could not be proven complete

The GEMM runs on 16-bit operands; the answer does not: the final ids and their order always come from fp32
cosines (the torch formula the reference calls), and a user only keeps the shortlist answer when the quantisation
error bound proves that no item outside the shortlist can enter its top-k by more than ``tie_tol`` -- the
tolerance below which the parity contract (and ``np.argsort`` in the reference) treats scores as tied.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _native as N
from . import ops

COS_EPS = 1e-6  # nn.CosineSimilarity(dim=1, eps=1e-6), src/metrics.py:58


@dataclass
class RecsConfig:
    elem: str = 'bf16'        # 16-bit operand type of the tensor-core GEMM: 'bf16' | 'fp16'
    parts: int = 2            # 1 = single product; 2 = hi/lo split, 3 products (hi.hi + lo.hi + hi.lo)
    shortlist: int = 16       # candidates kept per user by the GEMM epilogue (>= k, <= 32)
    center: bool = True       # subtract the mean normalised item row (ranking-invariant, shrinks the error bound)
    tie_tol: float = 1e-5     # score gap treated as a tie (north_star parity rule); 0 = strict
    acc_err: float = 1.5e-6   # allowance for fp32 accumulation error of the tensor-core sum, relative to |x||y|
    exact_only: bool = False  # skip the tensor-core path (brute-force fp32 kernel for every user)

    @property
    def elem_type(self) -> int:
        return {'bf16': N.ELEM_BF16, 'fp16': N.ELEM_FP16}[self.elem]

    def err_rel(self) -> float:
        """Worst-case |approx - exact| / (|x| |y|) of the quantised product scheme (unit roundoff u per operand:
        single product 2u + u^2; split scheme 3u^2 (1 + u)^2), plus the accumulation allowance."""
        u = 2.0 ** -9 if self.elem == 'bf16' else 2.0 ** -11
        q = (2 * u + u * u) if self.parts == 1 else 3 * u * u * (1 + u) ** 2
        return q + self.acc_err

    def err_abs(self, d: int) -> float:
        """fp16 only: lo halves and tiny values are subnormal (absolute spacing 2^-24): (|x|_1 + |y|_1) 2^-25."""
        return 0.0 if self.elem == 'bf16' else 2.0 * (d ** 0.5) * 2.0 ** -25


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
    """Item side of the scoring GEMM, prepared once per embedding table: quantised rows, centre, error stats."""

    def __init__(self, h_item: torch.Tensor, cfg: RecsConfig, item_id_base: int = 0):
        self.h_item = h_item.contiguous()
        self.cfg, self.item_id_base = cfg, int(item_id_base)
        self.n_items, self.d = h_item.shape
        self.tc = (not cfg.exact_only) and self.d <= 128 and self.n_items > 0
        self.center, self.items_q, self.stats = None, None, None
        if self.tc:
            self.d_pad = 64 if self.d <= 64 else 128
            if cfg.center:
                self.center = ops.colmean_normalized(self.h_item)
            self.items_q, self.stats = ops.score_prep(self.h_item, self.center, self.d_pad, cfg.parts, cfg.elem_type, True)


def recommend_topk(h_user: torch.Tensor, table: ScoringTable, k: int, bought: Optional[BoughtCSR] = None,
                   return_overflow: bool = False, mark=None, popularity: Optional[torch.Tensor] = None,
                   weight_popularity: float = 1.0):
    """Top-``k`` items of ``table`` for every row of ``h_user`` (``[n, d]`` fp32, CUDA): ``(ids int32 [n, k],
    scores fp32 [n, k])`` sorted by (score desc, id asc); ``-1`` / ``-inf`` pad rows with fewer than k candidates.
    ``bought`` rows must follow ``h_user`` rows. ids are global (``table.item_id_base`` added).
    ``mark(name)`` (optional) is called between stages -- bench.py records CUDA events with it."""
    mark = mark or (lambda name: None)
    cfg = table.cfg
    h_user = h_user.contiguous()
    n = h_user.shape[0]
    dev = h_user.device
    bptr, bids = (None, None) if bought is None else bought.on(dev)
    if bought is not None and bought.n_rows != n:
        raise ValueError('bought rows (%d) do not match user rows (%d)' % (bought.n_rows, n))
    if popularity is not None or not table.tc:
        # popularity re-rank (softmax over ALL items + w * popularity) runs on the exact fp32 kernel only
        ids, scores = ops.score_topk_exact(h_user, table.h_item, table.item_id_base, bptr, bids, k, COS_EPS,
                                           popularity=popularity, weight_popularity=weight_popularity)
        return (ids, scores, torch.zeros(1, dtype=torch.int32, device=dev)) if return_overflow else (ids, scores)
    shortlist = max(cfg.shortlist, k)
    if shortlist > 32:
        raise ValueError('k / shortlist above 32 is not supported by the fused top-k epilogue')
    users_q, _ = ops.score_prep(h_user, None, table.d_pad, cfg.parts, cfg.elem_type, False)
    mark('score_begin')
    sl_score, sl_id = ops.score_topk_tc(users_q, table.items_q, table.item_id_base, table.d_pad, cfg.parts,
                                        cfg.elem_type, bptr, bids, shortlist)
    mark('score_end')
    ids, scores, overflow, n_overflow = ops.rescore_topk(
        h_user, table.h_item, table.item_id_base, table.center, sl_score, sl_id, table.stats, cfg.err_rel(),
        cfg.err_abs(table.d), cfg.tie_tol, k, COS_EPS)
    # users whose shortlist could not be proven complete: exact fp32 pass (device-side count, no host sync)
    ops.score_topk_exact(h_user, table.h_item, table.item_id_base, bptr, bids, k, COS_EPS, user_list=overflow,
                         n_list=n_overflow, out_ids=ids, out_scores=scores)
    return (ids, scores, n_overflow) if return_overflow else (ids, scores)
