"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [group ...]      groups: embedding popularity sport forward metrics loss (default: all)

``generate(out_dir, groups)`` is what tests/test_golden_regen.py calls to prove that the committed fixtures are exactly
what this script writes today (every array compared bit for bit).

The reference's unmodified ``src/model.py`` (ConvModel, max_margin_loss), ``src/train/run.py::get_embeddings``
and ``src/metrics.py::{get_recs, create_already_bought}`` are imported from /root/reference and executed on CPU
over ``oracle/dgl_shim`` (DGL 0.5.2 is not installable here -- see the shim's docstring for what that pins
and what it does not). Inputs, weights and outputs of every case are written to ``<case>.npz``; the tests
compare the oracle (tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py) against them.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get('GNN_RECSYS_REFERENCE', '/root/reference')
sys.path[:0] = [os.path.join(ROOT, 'oracle', 'dgl_shim'), REFERENCE, ROOT]

import dgl  # noqa: E402  (the shim)
from src.model import ConvModel, max_margin_loss  # noqa: E402  (reference, verbatim)
from src.train.run import get_embeddings  # noqa: E402
from src.metrics import get_recs, create_already_bought, recs_to_metrics, get_metrics_at_k  # noqa: E402

import gnn_recsys_b200 as grb  # noqa: E402  (host-side containers only: synthetic data + block building)

OUT_DIR = HERE  # generate() redirects this
REL = [('user', 'buys', 'item'), ('item', 'bought-by', 'user'), ('user', 'clicks', 'item'), ('item', 'clicked-by', 'user')]


def tiny_data(n_users, n_items, n_edges, seed):
    """Synthetic graph plus the engineered edge cases of SURVEY.md 8c: an item and a user without
    in-edges, duplicate edges, a hub item, a user whose only edge is a click."""
    d = grb.make_graph(n_users, n_items, n_edges, seed)
    users, items, is_buy = d.users.copy(), d.items.copy(), d.is_buy.copy()
    lonely_item, lonely_user, hub = 3, 5, 7
    items[items == lonely_item] = hub                    # item 3 has no in-edges; item 7 becomes a hub
    users[users == lonely_user] = (lonely_user + 1) % n_users   # user 5 has no edges at all
    users[2:6], items[2:6], is_buy[2:6] = users[2], items[2], True   # 4 duplicate purchases
    d.users, d.items, d.is_buy = users, items, is_buy
    return d


def shim_graph(d, occurrence=False, seed=0):
    rel = d.relations()
    g = dgl.heterograph({c: (torch.from_numpy(s.astype(np.int64)), torch.from_numpy(t.astype(np.int64)))
                         for c, (s, t) in rel.items()}, {'user': d.n_users, 'item': d.n_items})
    g.nodes['user'].data['features'] = d.user_feat
    g.nodes['item'].data['features'] = d.item_feat
    occ = {}
    if occurrence:
        rng = np.random.default_rng(seed + 100)
        nb, nc = int(d.is_buy.sum()), int((~d.is_buy).sum())
        occ = {'buys': rng.integers(1, 5, nb), 'clicks': rng.integers(1, 5, nc)}
        occ['bought-by'], occ['clicked-by'] = occ['buys'], occ['clicks']
        for et, v in occ.items():
            g.edges[et].data['occurrence'] = torch.from_numpy(v.astype(np.int64))  # src/utils_data.py:304-315
    return g, occ


def to_shim_block(b):
    """product Block -> shim block (structure only; same local ids, CSR slot order)."""
    edges, ef = {}, {}
    for c, r in b.rels.items():
        dst = torch.repeat_interleave(torch.arange(r.n_dst), (r.indptr[1:] - r.indptr[:-1]).long())
        edges[c] = (r.indices.long(), dst)
        ef[c] = {} if r.weight is None else {'occurrence': r.weight}
    sf = {t: dict(f) for t, f in b._src_frames.items()}
    df = {t: dict(f) for t, f in b._dst_frames.items()}
    return dgl.DGLHeteroGraph(edges, b.num_src, b.num_dst, is_block=True, src_frames=sf, dst_frames=df, edge_frames=ef)


def save_case(name, meta, arrays):
    flat = {'meta': np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    for k, v in arrays.items():
        flat[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **flat)
    print('wrote %-34s %7.1f KB' % (name + '.npz', os.path.getsize(path) / 1024))


def edges_arrays(d, occ):
    out = {}
    for c, (s, t) in d.relations().items():
        out['edges/%s/src' % c[1]] = s.astype(np.int64)
        out['edges/%s/dst' % c[1]] = t.astype(np.int64)
    for et, v in occ.items():
        out['occurrence/%s' % et] = v.astype(np.int64)
    out['user_feat'], out['item_feat'] = d.user_feat, d.item_feat
    return out


def embedding_case(name, n_users, n_items, n_edges, seed, n_layers, hidden, out, aggregator, norm=True,
                   embedding_layer=True, hetero='sum', k=10, batched=None, sub_users=None):
    d = tiny_data(n_users, n_items, n_edges, seed)
    occ_on = aggregator.endswith('_edge')
    g, occ = shim_graph(d, occ_on, seed)
    torch.manual_seed(seed + 1)
    dim_dict = {'user': 2, 'item': 4, 'hidden': hidden, 'out': out}
    model = ConvModel(g, n_layers, dim_dict, norm, 0.0, aggregator, 'cos', hetero, embedding_layer)
    model.eval()
    conv_layers = n_layers - 1 if embedding_layer else n_layers
    uids = np.arange(n_users) if sub_users is None else np.asarray(sub_users)
    nids = {'user': torch.from_numpy(uids), 'item': torch.arange(n_items)}
    sampler = dgl.dataloading.MultiLayerFullNeighborSampler(conv_layers)
    bs = (n_users + n_items) if batched is None else batched
    torch.manual_seed(seed + 2)
    loader = dgl.dataloading.NodeDataLoader(g, nids, sampler, batch_size=bs, shuffle=batched is not None, drop_last=False)
    with torch.no_grad(), redirect_stdout(io.StringIO()):
        y = get_embeddings(g, out, model, loader, len(loader), False, None, embedding_layer)   # reference code
        bought_eids = g.out_edges(u=torch.from_numpy(uids), form='eid', etype='buys')          # main_inference.py:98
        bought = create_already_bought(g, bought_eids)                                         # reference code
        recs = get_recs(g, y, model, out, k, uids.tolist(), bought, remove_already_bought=True,
                        cuda=False, device=None, pred='cos', use_popularity=False)             # reference code
        recs_keep = get_recs(g, y, model, out, k, uids.tolist()[:8], bought, remove_already_bought=False)
    rec_arr = np.full((uids.size, k), -1, dtype=np.int64)
    for r, u in enumerate(uids.tolist()):
        rec_arr[r, :len(recs[u])] = np.asarray(recs[u], dtype=np.int64)
    keep_arr = np.stack([np.asarray(recs_keep[u][:k], dtype=np.int64) for u in uids.tolist()[:8]])
    meta = dict(name=name, n_users=n_users, n_items=n_items, n_layers=n_layers, hidden=hidden, out=out,
                aggregator=aggregator, norm=norm, embedding_layer=embedding_layer, hetero=hetero, k=k,
                batched=batched, seed=seed)
    arrays = edges_arrays(d, occ)
    arrays.update({'sd/' + kk: v for kk, v in model.state_dict().items()})
    arrays.update({'emb/user': y['user'], 'emb/item': y['item'], 'user_ids': uids, 'recs': rec_arr, 'recs_keep': keep_arr})
    save_case(name, meta, arrays)
    return d, g, model


def forward_case(name, n_users, n_items, n_edges, seed, hidden, out, fanouts, batch, neg_k, aggregator='mean', full=False):
    """Config-4 style training-step forward: sampled blocks + positive/negative edge scoring + loss. `_edge` aggregators:
    the graph carries edata['occurrence'] and the loader's blocks pick it up like DGL's do (src/model.py:174)."""
    d = tiny_data(n_users, n_items, n_edges, seed)
    g, occ = shim_graph(d, aggregator.endswith('_edge'), seed)
    pg = d.graph()
    for et, v in occ.items():
        pg.edges[et].data['occurrence'] = torch.from_numpy(v.astype(np.int64))
    torch.manual_seed(seed + 1)
    n_layers = len(fanouts) + 1
    model = ConvModel(g, n_layers, {'user': 2, 'item': 4, 'hidden': hidden, 'out': out}, True, 0.0, aggregator,
                      'cos', 'sum', True)
    model.eval()
    sampler = grb.MultiLayerFullNeighborSampler(len(fanouts)) if full else grb.MultiLayerNeighborSampler(fanouts)
    eids = {'buys': np.arange(pg.num_edges('buys')), 'clicks': np.arange(pg.num_edges('clicks'))}
    loader = grb.EdgeDataLoader(pg, eids, sampler, exclude='reverse_types',
                                reverse_etypes={'buys': 'bought-by', 'bought-by': 'buys', 'clicks': 'clicked-by',
                                                'clicked-by': 'clicks'},
                                negative_sampler=grb.negative_sampler.Uniform(neg_k), batch_size=batch, shuffle=True,
                                seed=seed + 2)
    _, pos_g, neg_g, blocks = next(iter(loader))
    sblocks = [to_shim_block(b) for b in blocks]
    sizes = {t: blocks[-1].num_dst[t] for t in pg.ntypes}

    def shim_pair(pgp):
        return dgl.DGLHeteroGraph({c: tuple(torch.from_numpy(np.asarray(a, dtype=np.int64)) for a in pgp.edge_arrays(c))
                                   for c in pgp.canonical_etypes}, sizes, sizes)
    with torch.no_grad():
        feats = sblocks[0].srcdata['features']
        h, pos, neg = model(sblocks, feats, shim_pair(pos_g), shim_pair(neg_g), True)          # reference code
        loss = max_margin_loss(pos, neg, 0.266, neg_k)                                          # reference code
    arrays = {'sd/' + kk: v for kk, v in model.state_dict().items()}
    for li, b in enumerate(blocks):
        for t in pg.ntypes:
            arrays['block%d/nsrc/%s' % (li, t)] = np.int64(b.num_src[t])
            arrays['block%d/ndst/%s' % (li, t)] = np.int64(b.num_dst[t])
            arrays['block%d/srcid/%s' % (li, t)] = b.srcnodes[t].data[grb.NID]
        for c, r in b.rels.items():
            arrays['block%d/indptr/%s' % (li, c[1])] = r.indptr
            arrays['block%d/indices/%s' % (li, c[1])] = r.indices
            if r.weight is not None:
                arrays['block%d/weight/%s' % (li, c[1])] = r.weight
    for t in pg.ntypes:
        arrays['feat/' + t] = blocks[0].srcnodes[t].data['features']
        arrays['h/' + t] = h[t]
    for c in pg.canonical_etypes:
        for nm, gg, sc in (('pos', pos_g, pos), ('neg', neg_g, neg)):
            s, t = gg.edge_arrays(c)
            arrays['%s/%s/src' % (nm, c[1])], arrays['%s/%s/dst' % (nm, c[1])] = s, t
            arrays['%s/%s/score' % (nm, c[1])] = sc[c]
    arrays['loss'] = loss
    meta = dict(name=name, n_layers=n_layers, hidden=hidden, out=out, aggregator=aggregator, neg_k=neg_k, delta=0.266,
                n_blocks=len(blocks), seed=seed)
    save_case(name, meta, arrays)


def main_embedding():
    T = dict(n_users=50, n_items=20, n_edges=300)
    for agg in ('mean', 'mean_nn', 'pool_nn', 'mean_edge', 'pool_nn_edge'):
        embedding_case('tiny_%s' % agg, seed=3, n_layers=3, hidden=16, out=8, aggregator=agg, **T)
    embedding_case('tiny_mean_nonorm', seed=4, n_layers=2, hidden=16, out=8, aggregator='mean', norm=False, **T)
    embedding_case('tiny_mean_noembed', seed=5, n_layers=2, hidden=16, out=8, aggregator='mean', embedding_layer=False, **T)
    embedding_case('tiny_pool_hetero_max', seed=6, n_layers=3, hidden=16, out=8, aggregator='pool_nn', hetero='max', **T)
    embedding_case('tiny_mean_hetero_mean', seed=7, n_layers=2, hidden=16, out=8, aggregator='mean', hetero='mean', **T)
    # the reference's real batching (128 nodes, shuffle) must equal the one-block-per-layer pass for seeded rows
    embedding_case('tiny_mean_batched', seed=8, n_layers=3, hidden=16, out=8, aggregator='mean', batched=16,
                   sub_users=list(range(0, 50, 2)), **T)
    # model-sized dims of configs c1/c2 (2-layer mean 128/128) and c3 (3-layer pool_nn hidden 256)
    embedding_case('small_mean_128', 400, 150, 6000, seed=9, n_layers=2, hidden=128, out=128, aggregator='mean')
    embedding_case('small_pool_256', 200, 80, 3000, seed=10, n_layers=3, hidden=256, out=128, aggregator='pool_nn')
    # the other (hidden, out) presets of the reference's search space (main.py:86-87): Very Small, Small, Very Large
    embedding_case('preset_64_32', 120, 50, 1500, seed=15, n_layers=3, hidden=64, out=32, aggregator='mean')
    embedding_case('preset_192_96', 120, 50, 1500, seed=16, n_layers=3, hidden=192, out=96, aggregator='pool_nn')
    embedding_case('preset_512_256', 100, 40, 1200, seed=17, n_layers=2, hidden=512, out=256, aggregator='mean')



def popularity_case(base_name, weight, seed):
    """use_popularity branch of the reference's get_recs (src/metrics.py:69-72) on the embeddings of an existing
    fixture: item popularity = share of purchases (like src/builder.py:472-491 a [I, 1] float tensor)."""
    z = np.load(os.path.join(OUT_DIR, base_name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    n_users, n_items, k = meta['n_users'], meta['n_items'], meta['k']
    rel = {c: (torch.from_numpy(z['edges/%s/src' % c[1]]), torch.from_numpy(z['edges/%s/dst' % c[1]])) for c in REL}
    g = dgl.heterograph(rel, {'user': n_users, 'item': n_items})
    buys_items = z['edges/buys/dst']
    pop = np.bincount(buys_items, minlength=n_items).astype(np.float64)
    pop = (pop / max(pop.sum(), 1.0)).astype(np.float32).reshape(-1, 1)
    g.nodes['item'].data['popularity'] = torch.from_numpy(pop)
    g.nodes['user'].data['popularity'] = torch.zeros(n_users, 1)
    y = {'user': torch.from_numpy(z['emb/user']), 'item': torch.from_numpy(z['emb/item'])}
    uids = z['user_ids']
    with torch.no_grad(), redirect_stdout(io.StringIO()):
        bought = create_already_bought(g, g.out_edges(u=torch.from_numpy(uids), form='eid', etype='buys'))
        recs = get_recs(g, y, None, meta['out'], k, uids.tolist(), bought, remove_already_bought=True, cuda=False,
                        device=None, pred='cos', use_popularity=True, weight_popularity=weight)   # reference code
    rec_arr = np.full((uids.size, k), -1, dtype=np.int64)
    for r, u in enumerate(uids.tolist()):
        rec_arr[r, :len(recs[u])] = np.asarray(recs[u], dtype=np.int64)
    save_case(base_name + '_pop', dict(base=base_name, weight=weight, k=k), {'popularity': pop, 'recs_pop': rec_arr})


def main_popularity():
    popularity_case('tiny_mean', 1.0, 0)
    popularity_case('small_mean_128', 0.5, 1)




SPORT_REL = [('item', 'utilized-for', 'sport'), ('sport', 'utilizes', 'item'), ('user', 'practices', 'sport'),
             ('sport', 'practiced-by', 'user'), ('sport', 'belongs-to', 'sport'), ('sport', 'includes', 'sport')]


def sport_case(name, aggregator, seed, n_layers=3, hidden=16, out=8, n_users=50, n_items=20, n_edges=300, n_sports=6):
    """The full 10-relation schema with a third node type (src/utils_data.py:204-238, include_sport=True), embedded
    the way main_inference.py does it: only users and items are seeded, so the 'sport' table stays zero
    (src/train/run.py:329-333). `_edge` aggregators multiply by `occurrence` on user-item relations only and fall
    back to copy_src on every relation touching 'sport' (src/model.py:171-208)."""
    d = tiny_data(n_users, n_items, n_edges, seed)
    rng = np.random.default_rng(seed + 50)
    rel = dict(d.relations())
    i2s = (np.arange(n_items), rng.integers(0, n_sports, n_items))
    u2s = (rng.integers(0, n_users, 80), rng.integers(0, n_sports, 80))
    s2g = (np.arange(n_sports), (np.arange(n_sports) + 1) % 3)
    rel.update({SPORT_REL[0]: i2s, SPORT_REL[1]: (i2s[1], i2s[0]), SPORT_REL[2]: u2s, SPORT_REL[3]: (u2s[1], u2s[0]),
                SPORT_REL[4]: s2g, SPORT_REL[5]: (s2g[1], s2g[0])})
    num = {'user': n_users, 'item': n_items, 'sport': n_sports}
    g = dgl.heterograph({c: (torch.from_numpy(np.asarray(s, dtype=np.int64)), torch.from_numpy(np.asarray(t, dtype=np.int64)))
                         for c, (s, t) in rel.items()}, num)
    sport_feat = torch.from_numpy(rng.standard_normal((n_sports, 6)).astype(np.float32))
    g.nodes['user'].data['features'] = d.user_feat
    g.nodes['item'].data['features'] = d.item_feat
    g.nodes['sport'].data['features'] = sport_feat
    occ = {}
    if aggregator.endswith('_edge'):
        nb, nc = int(d.is_buy.sum()), int((~d.is_buy).sum())
        occ = {'buys': rng.integers(1, 5, nb), 'clicks': rng.integers(1, 5, nc)}
        occ['bought-by'], occ['clicked-by'] = occ['buys'], occ['clicks']
        for et, v in occ.items():
            g.edges[et].data['occurrence'] = torch.from_numpy(v.astype(np.int64))
    torch.manual_seed(seed + 1)
    dim_dict = {'user': 2, 'item': 4, 'sport': 6, 'hidden': hidden, 'out': out}
    model = ConvModel(g, n_layers, dim_dict, True, 0.0, aggregator, 'cos', 'sum', True)
    model.eval()
    nids = {'user': torch.arange(n_users), 'item': torch.arange(n_items)}           # main_inference.py:125
    sampler = dgl.dataloading.MultiLayerFullNeighborSampler(n_layers - 1)
    loader = dgl.dataloading.NodeDataLoader(g, nids, sampler, batch_size=n_users + n_items, shuffle=False, drop_last=False)
    with torch.no_grad(), redirect_stdout(io.StringIO()):
        y = get_embeddings(g, out, model, loader, len(loader), False, None, True)   # reference code
    assert float(y['sport'].abs().max()) == 0.0
    arrays = {'sd/' + kk: v for kk, v in model.state_dict().items()}
    for c, (s, t) in rel.items():
        arrays['edges/%s/src' % c[1]] = np.asarray(s, dtype=np.int64)
        arrays['edges/%s/dst' % c[1]] = np.asarray(t, dtype=np.int64)
    for et, v in occ.items():
        arrays['occurrence/%s' % et] = v.astype(np.int64)
    arrays.update({'feat/user': d.user_feat, 'feat/item': d.item_feat, 'feat/sport': sport_feat,
                   'emb/user': y['user'], 'emb/item': y['item'], 'emb/sport': y['sport']})
    meta = dict(name=name, num=num, rels=[list(c) for c in rel], n_layers=n_layers, hidden=hidden, out=out,
                aggregator=aggregator, dims=dim_dict, seed=seed)
    save_case(name, meta, arrays)


def main_sport():
    sport_case('sport_mean_edge', 'mean_edge', 21)
    sport_case('sport_pool_nn', 'pool_nn', 22)


def main_forward():
    """Training-step forwards (config 4 shape), incl. dims that take the fused tensor-core ConvLayer kernel on SAMPLED
    blocks (n_src != n_dst, destination prefix), the full-neighbour sampler and a pool_nn stack with fc_preagg."""
    forward_case('fwd_fanout_mean', 300, 120, 5000, seed=11, hidden=32, out=16, fanouts=[10, 10], batch=64, neg_k=20)
    forward_case('fwd_fanout_mean_128', 400, 150, 6000, seed=12, hidden=128, out=128, fanouts=[10, 10], batch=128, neg_k=50)
    forward_case('fwd_full_pool_nn', 300, 120, 5000, seed=13, hidden=128, out=64, fanouts=[0, 0], batch=64, neg_k=20,
                 aggregator='pool_nn', full=True)
    forward_case('fwd_fanout_mean_edge', 300, 120, 5000, seed=14, hidden=32, out=16, fanouts=[10, 10], batch=64, neg_k=20,
                 aggregator='mean_edge')


def metrics_case(base_name, seed, k_big):
    """recs_to_metrics / get_metrics_at_k of the reference (src/metrics.py:81-134) on the embeddings of an existing
    fixture. Ground truth = random (user, item) pairs with duplicate pairs and users without any recommendation hit;
    `recs` come from the reference's own get_recs (k from the base case, and k_big > 32 for the large-k path)."""
    z = np.load(os.path.join(OUT_DIR, base_name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    n_users, n_items, k = meta['n_users'], meta['n_items'], meta['k']
    rel = {c: (torch.from_numpy(z['edges/%s/src' % c[1]]), torch.from_numpy(z['edges/%s/dst' % c[1]])) for c in REL}
    g = dgl.heterograph(rel, {'user': n_users, 'item': n_items})
    y = {'user': torch.from_numpy(z['emb/user']), 'item': torch.from_numpy(z['emb/item'])}
    rng = np.random.default_rng(seed)
    n_gt = 3 * n_users
    gt_users = rng.integers(0, n_users, n_gt)
    gt_users[gt_users % 7 == 3] = 0                      # some users absent from the ground truth, user 0 heavy
    gt_items = rng.integers(0, n_items, n_gt)
    gt_users[5:9], gt_items[5:9] = gt_users[5], gt_items[5]   # duplicate ground-truth pairs (counted like the reference)
    uids = np.unique(gt_users)
    bought_eids = g.out_edges(u=torch.from_numpy(uids), form='eid', etype='buys')
    out = {'gt_users': gt_users.astype(np.int64), 'gt_items': gt_items.astype(np.int64), 'bought_eids': bought_eids}
    with torch.no_grad(), redirect_stdout(io.StringIO()):
        for kk in (k, k_big):
            for rm in (True, False):
                p, r, c = get_metrics_at_k(y, g, None, meta['out'], (gt_users, gt_items), bought_eids, kk, rm, False,
                                           None, 'cos', False, 1)                                   # reference code
                out['metrics/k%d/remove%d' % (kk, int(rm))] = np.array([p, r, c], dtype=np.float64)
        # recs_to_metrics alone on hand-made, ragged recommendation lists (incl. an empty one)
        recs = {int(u): rng.choice(n_items, int(rng.integers(0, 6)), replace=False).tolist() for u in uids.tolist()}
        recs[int(uids[0])] = []
        from src.metrics import create_ground_truth
        p, r, c = recs_to_metrics(recs, create_ground_truth(gt_users, gt_items), g)                 # reference code
    lens = np.array([len(recs[int(u)]) for u in uids.tolist()], dtype=np.int64)
    out['ragged/users'], out['ragged/lens'] = uids.astype(np.int64), lens
    out['ragged/items'] = np.array([i for u in uids.tolist() for i in recs[int(u)]], dtype=np.int64)
    out['ragged/metrics'] = np.array([p, r, c], dtype=np.float64)
    save_case(base_name + '_metrics', dict(base=base_name, k=k, k_big=k_big, seed=seed), out)


def main_metrics():
    metrics_case('tiny_mean', 31, 13)          # 20 items: k_big = 13 leaves rows with fewer than k candidates
    metrics_case('small_mean_128', 32, 40)     # k = 40 > 32


def loss_case(base_name, seed):
    """max_margin_loss of the reference (src/model.py:473-533) with remove_false_negative and / or use_recency on the
    positive / negative scores of an existing forward fixture. negative_mask = has_edges_between as float
    (src/train/run.py:100-101), recency only for 'buys' (the KeyError branch covers the etypes without it)."""
    z = np.load(os.path.join(OUT_DIR, base_name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    neg_k, delta = meta['neg_k'], meta['delta']
    rng = np.random.default_rng(seed)
    pos, neg, mask, rec = {}, {}, {}, {}
    for c in REL:
        key = 'pos/%s/score' % c[1]
        if key not in z.files or z[key].shape[0] == 0:
            continue
        pos[c], neg[c] = torch.from_numpy(z[key]), torch.from_numpy(z['neg/%s/score' % c[1]])
        mask[c] = torch.from_numpy((rng.random(neg[c].shape[0]) < 0.15).astype(np.float32))
    rec[REL[0]] = torch.from_numpy((1.0 + 9.0 * rng.random(pos[REL[0]].shape[0])).astype(np.float32))
    out = {}
    for c in pos:
        out['mask/%s' % c[1]] = mask[c]
    out['recency/buys'] = rec[REL[0]]
    for rfn in (False, True):
        for ur in (False, True):
            with torch.no_grad():
                l = max_margin_loss(pos, neg, delta, neg_k, use_recency=ur, recency_scores=rec,
                                    remove_false_negative=rfn, negative_mask=mask)                   # reference code
            out['loss/mask%d/recency%d' % (int(rfn), int(ur))] = l
    save_case(base_name + '_loss', dict(base=base_name, neg_k=neg_k, delta=delta, seed=seed), out)


def main_loss():
    loss_case('fwd_fanout_mean', 41)
    loss_case('fwd_fanout_mean_128', 42)


GROUPS = {'embedding': main_embedding, 'popularity': main_popularity, 'sport': main_sport, 'forward': main_forward,
          'metrics': main_metrics, 'loss': main_loss}   # in dependency order (popularity / metrics / loss read base cases)


def generate(out_dir=HERE, groups=None):
    global OUT_DIR
    OUT_DIR = out_dir
    os.makedirs(out_dir, exist_ok=True)
    try:
        for name, fn in GROUPS.items():
            if groups is None or name in groups:
                fn()
    finally:
        OUT_DIR = HERE


if __name__ == '__main__':
    generate(HERE, sys.argv[1:] or None)
