"""``get_recs`` / ``create_already_bought`` -- the reference's ``src/metrics.py:19-78`` on the B200 kernels.

``get_recs`` keeps the reference signature and return value (``{user id: list of k item ids}``); the per-user
Python loop (repeat the user row, cosine against every item, D2H, ``np.argsort``, Python filter) is replaced by
the device pipeline of ``recs.py``. ``get_recs_tensor`` is the same computation without the Python dict at the
end -- what a 10M-user job should call.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Optional

import numpy as np
import torch

from .recs import BoughtCSR, RecsConfig, ScoringTable, recommend_topk


def create_already_bought(g, bought_eids, etype='buys'):
    """Dictionary user id -> item ids the user already bought (reference ``src/metrics.py:19-28``)."""
    users_train, items_train = g.find_edges(bought_eids, etype=etype)
    already_bought_dict = defaultdict(list)
    for key, val in zip(np.asarray(users_train).tolist(), np.asarray(items_train).tolist()):
        already_bought_dict[key].append(val)
    return already_bought_dict


def create_already_bought_csr(g, bought_eids, etype='buys') -> BoughtCSR:
    """Same content as ``create_already_bought`` as a ``BoughtCSR`` (vectorised; rows = user ids)."""
    users_train, items_train = g.find_edges(bought_eids, etype=etype)
    c = g.to_canonical_etype(etype)
    return BoughtCSR.from_edges(np.asarray(users_train), np.asarray(items_train), g.num_nodes(c[0]))


def _device_of(h, device):
    if device is not None:
        return torch.device(device)
    if h['item'].is_cuda:
        return h['item'].device
    return torch.device('cuda', torch.cuda.current_device())  # raises without a GPU: no CPU fallback


@torch.no_grad()
def get_recs_tensor(g, h, k, user_ids, already_bought=None, remove_already_bought=True, device=None,
                    config: Optional[RecsConfig] = None, table: Optional[ScoringTable] = None,
                    return_scores: bool = False, use_popularity: bool = False, weight_popularity=1):
    """``[len(user_ids), k]`` int32 item ids on the device (``-1`` where fewer than k items qualify)."""
    dev = _device_of(h, device)
    cfg = config or RecsConfig()
    h_item = h['item'].to(dev, torch.float32)
    assert h_item.shape[0] == g.num_nodes('item')
    if table is None:
        table = ScoringTable(h_item, cfg)
    hu_all = h['user'].to(dev, torch.float32)
    uid = np.asarray(user_ids, dtype=np.int64).reshape(-1)
    if uid.size == hu_all.shape[0] and np.array_equal(uid, np.arange(uid.size)):
        hu = hu_all
    else:
        hu = hu_all[torch.from_numpy(uid).to(dev)]
    bought = None
    if remove_already_bought and already_bought is not None:
        if isinstance(already_bought, BoughtCSR):
            bought = already_bought.select(uid)
        else:
            bought = BoughtCSR.from_dict(already_bought, uid.tolist())
    pop = None
    if use_popularity:  # g.ndata['popularity']['item'] like the reference (src/metrics.py:71)
        pop = g.nodes['item'].data['popularity'].to(dev, torch.float32).reshape(-1).contiguous()
    ids, scores = recommend_topk(hu, table, k, bought, popularity=pop, weight_popularity=float(weight_popularity))
    return (ids, scores) if return_scores else ids


def get_recs(g, h, model, embed_dim, k, user_ids, already_bought_dict, remove_already_bought=True, cuda=False,
             device=None, pred: str = 'cos', use_popularity: bool = False, weight_popularity=1):
    """Computes K recommendations for all users, given hidden states and what they already bought."""
    if pred == 'nn':
        raise NotImplementedError("pred='nn' (MLP scorer) is outside the accelerated hot path (DESIGN.md, scope)")
    if pred != 'cos':
        raise KeyError(f'Prediction function {pred} not recognized.')
    print('Computing recommendations on {} users, for {} items'.format(len(user_ids), g.num_nodes('item')))
    ids = get_recs_tensor(g, h, k, user_ids, already_bought_dict, remove_already_bought, device,
                          use_popularity=use_popularity, weight_popularity=weight_popularity).cpu().numpy()
    ids = ids.astype(np.int64)
    if not remove_already_bought:  # the reference returns the ndarray slice order[:k] on this branch
        return {user: ids[r][ids[r] >= 0] for r, user in enumerate(user_ids)}
    rows = ids.tolist()
    short = np.nonzero((ids < 0).any(axis=1))[0]  # users with fewer than k candidates (rare): drop the padding
    for r in short.tolist():
        rows[r] = [i for i in rows[r] if i >= 0]
    return dict(zip(list(user_ids), rows))


def create_ground_truth(users, items):
    """Dictionary user id -> item ids the user actually bought (reference ``src/metrics.py:8-16``)."""
    ground_truth_dict = defaultdict(list)
    for key, val in zip(np.asarray(users).tolist(), np.asarray(items).tolist()):
        ground_truth_dict[key].append(val)
    return ground_truth_dict


def metrics_from_tensor(ids: torch.Tensor, truth: BoughtCSR, n_items: int):
    """precision, recall, coverage of an ``[n, k]`` id table against a ground-truth CSR whose rows follow the rows
    of ``ids`` -- ``recs_to_metrics`` (reference ``src/metrics.py:81-107``) computed on the device."""
    from . import ops
    tptr, tids = truth.on(ids.device)
    c = ops.metrics_at_k(ids.contiguous(), tptr, tids, n_items).cpu().tolist()
    return c[1] / c[0], c[3] / c[2], c[4] / n_items


def recs_to_metrics(recs, ground_truth_dict, g):
    """Given the recommendations and the ground truth, computes precision, recall & coverage (dict API of the
    reference; the counting runs on the device)."""
    users = list(recs.keys())
    k = max((len(v) for v in recs.values()), default=1)
    ids = np.full((len(users), max(k, 1)), -1, dtype=np.int32)
    for r, u in enumerate(users):
        row = np.asarray(recs[u], dtype=np.int32)
        ids[r, :row.size] = row
    truth = BoughtCSR.from_dict(ground_truth_dict, users)
    dev = torch.device('cuda', torch.cuda.current_device())
    return metrics_from_tensor(torch.from_numpy(ids).to(dev), truth, g.num_nodes('item'))


def get_metrics_at_k(h, g, model, embed_dim, ground_truth, bought_eids, k, remove_already_bought=True, cuda=False,
                     device=None, pred='cos', use_popularity=False, weight_popularity=1):
    """create already_bought & ground truth, get recs and compute metrics (reference ``src/metrics.py:110-134``),
    without the Python dicts in between."""
    if pred == 'nn':
        raise NotImplementedError("pred='nn' (MLP scorer) is outside the accelerated hot path (DESIGN.md, scope)")
    if pred != 'cos':
        raise KeyError(f'Prediction function {pred} not recognized.')
    users, items = ground_truth
    users, items = np.asarray(users), np.asarray(items)
    user_ids = np.unique(users)
    n_users = g.num_nodes('user')
    bought = create_already_bought_csr(g, bought_eids) if remove_already_bought else None
    truth = BoughtCSR.from_edges(users, items, n_users).select(user_ids)
    ids = get_recs_tensor(g, h, k, user_ids, bought, remove_already_bought, device, use_popularity=use_popularity,
                          weight_popularity=weight_popularity)
    return metrics_from_tensor(ids, truth, g.num_nodes('item'))
