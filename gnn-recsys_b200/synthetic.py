"""Seeded synthetic user-item graphs shaped like the reference's data.

The reference needs proprietary Decathlon CSVs (``README.md:16``); benchmarks and parity tests use this
generator instead: a click / purchase multigraph with power-law popularity.

  * item of each edge  ~ Zipf(1.0) over a fixed random permutation of item ids
  * user of each edge  ~ p(rank) ∝ rank^-0.5 over a fixed random permutation of user ids
  * multi-edges are kept (reference default ``duplicates='keep_all'``, ``src/builder.py:275``)
  * each edge is a purchase with probability 0.2, else a click (``discern_clicks``); the reverse relations
    ``bought-by`` / ``clicked-by`` are the same edges swapped (``src/utils_data.py:205-214``)
  * features: user ``[U, 2]`` one-hot gender flags (``src/builder.py:426-434``), item ``[I, 4]`` Bernoulli(0.3)
    flags (``src/builder.py:444-453``), fp32
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np
import torch

from .graph import HeteroGraph

CONFIGS = {
    # name: (users, items, edges, n_layers, aggregator, hidden, out)   -- BASELINE.json configs[0..4]
    'c1': (10_000, 5_000, 200_000, 2, 'mean', 128, 128),
    'c2': (1_000_000, 200_000, 50_000_000, 2, 'mean', 128, 128),
    'c3': (5_000_000, 500_000, 200_000_000, 3, 'pool_nn', 256, 128),
    'c5': (10_000_000, 1_000_000, 500_000_000, 2, 'mean', 128, 128),
}


@dataclass
class SyntheticData:
    n_users: int
    n_items: int
    users: np.ndarray        # int32 [E] edge sources (user ids), edge order = generation order
    items: np.ndarray        # int32 [E] edge destinations (item ids)
    is_buy: np.ndarray       # bool  [E]
    user_feat: torch.Tensor  # fp32 [U, 2]
    item_feat: torch.Tensor  # fp32 [I, 4]

    def relations(self) -> Dict[Tuple[str, str, str], Tuple[np.ndarray, np.ndarray]]:
        b, c = self.is_buy, ~self.is_buy
        ub, ib, uc, ic = self.users[b], self.items[b], self.users[c], self.items[c]
        return {('user', 'buys', 'item'): (ub, ib), ('item', 'bought-by', 'user'): (ib, ub),
                ('user', 'clicks', 'item'): (uc, ic), ('item', 'clicked-by', 'user'): (ic, uc)}

    def graph(self) -> HeteroGraph:
        g = HeteroGraph(self.relations(), {'user': self.n_users, 'item': self.n_items})
        g.nodes['user'].data['features'] = self.user_feat
        g.nodes['item'].data['features'] = self.item_feat
        return g


def _power_law_draw(rng, n: int, size: int, alpha: float) -> np.ndarray:
    w = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    rank = np.searchsorted(cdf, rng.random(size), side='right')
    np.minimum(rank, n - 1, out=rank)
    perm = rng.permutation(n)
    return perm[rank].astype(np.int32)


def make_graph(n_users: int, n_items: int, n_edges: int, seed: int = 0) -> SyntheticData:
    rng = np.random.default_rng(seed)
    items = _power_law_draw(rng, n_items, n_edges, 1.0)
    users = _power_law_draw(rng, n_users, n_edges, 0.5)
    is_buy = rng.random(n_edges) < 0.2
    if n_edges >= 2:  # make the node counts exact: the largest ids appear, once as a buy and once as a click
        users[0], items[0], is_buy[0] = n_users - 1, n_items - 1, True
        users[1], items[1], is_buy[1] = n_users - 1, n_items - 1, False
    gender = rng.integers(0, 2, size=n_users)
    user_feat = np.zeros((n_users, 2), dtype=np.float32)
    user_feat[np.arange(n_users), gender] = 1.0
    item_feat = (rng.random((n_items, 4)) < 0.3).astype(np.float32)
    return SyntheticData(n_users, n_items, users, items, is_buy,
                         torch.from_numpy(user_feat), torch.from_numpy(item_feat))


def make_graph_device(n_users: int, n_items: int, n_edges: int, seed: int = 0, device='cuda') -> SyntheticData:
    """Same distributions as ``make_graph`` drawn with torch on ``device`` (seconds instead of minutes at 50M+
    edges; a different random stream, so use ``make_graph`` wherever a fixture must be reproducible on the CPU)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    def draw(n, alpha):
        w = torch.arange(1, n + 1, dtype=torch.float64, device=device) ** (-alpha)
        cdf = torch.cumsum(w, 0)
        cdf /= cdf[-1].clone()
        r = torch.rand(n_edges, dtype=torch.float64, device=device, generator=gen)
        rank = torch.searchsorted(cdf, r, right=True).clamp_(max=n - 1)
        perm = torch.randperm(n, device=device, generator=gen)
        return perm[rank].to(torch.int32)

    items, users = draw(n_items, 1.0), draw(n_users, 0.5)
    is_buy = torch.rand(n_edges, device=device, generator=gen) < 0.2
    if n_edges >= 2:
        users[0], items[0], is_buy[0] = n_users - 1, n_items - 1, True
        users[1], items[1], is_buy[1] = n_users - 1, n_items - 1, False
    gender = torch.randint(0, 2, (n_users,), device=device, generator=gen)
    user_feat = torch.zeros((n_users, 2), dtype=torch.float32, device=device)
    user_feat[torch.arange(n_users, device=device), gender] = 1.0
    item_feat = (torch.rand((n_items, 4), device=device, generator=gen) < 0.3).to(torch.float32)
    return SyntheticData(n_users, n_items, users.cpu().numpy(), items.cpu().numpy(), is_buy.cpu().numpy(),
                         user_feat.cpu(), item_feat.cpu())
