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
                    return_scores: bool = False):
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
    ids, scores = recommend_topk(hu, table, k, bought)
    return (ids, scores) if return_scores else ids


def get_recs(g, h, model, embed_dim, k, user_ids, already_bought_dict, remove_already_bought=True, cuda=False,
             device=None, pred: str = 'cos', use_popularity: bool = False, weight_popularity=1):
    """Computes K recommendations for all users, given hidden states and what they already bought."""
    if pred == 'nn':
        raise NotImplementedError("pred='nn' (MLP scorer) is outside the accelerated hot path (DESIGN.md, scope)")
    if pred != 'cos':
        raise KeyError(f'Prediction function {pred} not recognized.')
    if use_popularity:
        raise NotImplementedError('use_popularity re-ranking is not implemented yet (DESIGN.md, next)')
    print('Computing recommendations on {} users, for {} items'.format(len(user_ids), g.num_nodes('item')))
    ids = get_recs_tensor(g, h, k, user_ids, already_bought_dict, remove_already_bought, device).cpu().numpy()
    ids = ids.astype(np.int64)
    recs = {}
    for r, user in enumerate(user_ids):
        row = ids[r]
        row = row[row >= 0]
        recs[user] = list(row) if remove_already_bought else row
    return recs
