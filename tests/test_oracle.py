"""The straight-line oracle against the golden fixtures produced by the reference's own code."""
import numpy as np
import pytest
import torch

from helpers import EMBED_CASES, load_case, state_dict, case_relations, case_occurrence, assert_topk_equivalent, RELS
from oracle import straightline as O


def full_blocks(meta, z):
    rel, occ = case_relations(z), case_occurrence(z)
    num = {'user': meta['n_users'], 'item': meta['n_items']}
    blk = O.block_from_coo(num, num, {c: (s, d, occ.get(c)) for c, (s, d) in rel.items()})
    n_conv = meta['n_layers'] - 1 if meta['embedding_layer'] else meta['n_layers']
    return num, [blk] * n_conv


@pytest.mark.parametrize('name', EMBED_CASES)
def test_embeddings_match_reference(name):
    meta, z = load_case(name)
    num, blocks = full_blocks(meta, z)
    feats = {'user': torch.from_numpy(z['user_feat']), 'item': torch.from_numpy(z['item_feat'])}
    seeds = {'user': z['user_ids'], 'item': np.arange(meta['n_items'])}
    y = O.get_embeddings_full(num, blocks, feats, state_dict(z), meta['out'], seeds, meta['aggregator'], meta['norm'],
                              meta['hetero'], meta['embedding_layer'])
    for t in ('user', 'item'):
        np.testing.assert_allclose(y[t].numpy(), z['emb/' + t], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('name', EMBED_CASES)
def test_recs_match_reference(name):
    meta, z = load_case(name)
    hu, hi = torch.from_numpy(z['emb/user']), torch.from_numpy(z['emb/item'])
    buys = case_relations(z)[('user', 'buys', 'item')]
    bought = O.create_already_bought(buys[0], buys[1])
    uids = z['user_ids'].tolist()
    recs = O.get_recs(hu, hi, meta['k'], uids, bought)
    got = np.full((len(uids), meta['k']), -1, dtype=np.int64)
    for r, u in enumerate(uids):
        got[r, :len(recs[u])] = recs[u]
    scores = O.get_recs_scores(hu, hi, uids).numpy()
    assert_topk_equivalent(got, z['recs'], scores, meta['k'])
    # vectorised variant (the "fair" CPU baseline) gives the same answer up to ties
    order = np.lexsort((buys[1], buys[0]))
    indptr = np.zeros(meta['n_users'] + 1, dtype=np.int64)
    np.cumsum(np.bincount(buys[0], minlength=meta['n_users']), out=indptr[1:])
    vec = O.get_recs_vectorised(hu, hi, meta['k'], uids, indptr, buys[1][order])
    assert_topk_equivalent(vec, z['recs'], scores, meta['k'])


@pytest.mark.parametrize('name', ['fwd_fanout_mean', 'fwd_fanout_mean_128', 'fwd_full_pool_nn', 'fwd_fanout_mean_edge'])
def test_forward_scores_and_loss_match_reference(name):
    meta, z = load_case(name)
    blocks = []
    for li in range(meta['n_blocks']):
        ns = {t: int(z['block%d/nsrc/%s' % (li, t)]) for t in ('user', 'item')}
        nd = {t: int(z['block%d/ndst/%s' % (li, t)]) for t in ('user', 'item')}
        rels = {}
        for c in [('item', 'bought-by', 'user'), ('item', 'clicked-by', 'user'), ('user', 'buys', 'item'), ('user', 'clicks', 'item')]:
            indptr, indices = z['block%d/indptr/%s' % (li, c[1])], z['block%d/indices/%s' % (li, c[1])]
            dst = np.repeat(np.arange(indptr.size - 1), np.diff(indptr))
            wkey = 'block%d/weight/%s' % (li, c[1])
            rels[c] = (indices.astype(np.int64), dst, z[wkey] if wkey in z.files else None)
        blocks.append(O.block_from_coo(ns, nd, rels))
    feats = {t: torch.from_numpy(z['feat/' + t]) for t in ('user', 'item')}
    cets = [('item', 'bought-by', 'user'), ('item', 'clicked-by', 'user'), ('user', 'buys', 'item'), ('user', 'clicks', 'item')]
    pos = {c: (z['pos/%s/src' % c[1]], z['pos/%s/dst' % c[1]]) for c in cets}
    neg = {c: (z['neg/%s/src' % c[1]], z['neg/%s/dst' % c[1]]) for c in cets}
    h, ps, ns_ = O.model_forward(blocks, feats, pos, neg, state_dict(z), meta['aggregator'])
    for t in ('user', 'item'):
        np.testing.assert_allclose(h[t].numpy(), z['h/' + t], rtol=1e-4, atol=1e-5)
    for c in cets:
        np.testing.assert_allclose(ps[c].numpy(), z['pos/%s/score' % c[1]], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(ns_[c].numpy(), z['neg/%s/score' % c[1]], rtol=1e-4, atol=1e-5)
    loss = O.max_margin_loss(ps, ns_, meta['delta'], meta['neg_k'])
    np.testing.assert_allclose(float(loss), float(z['loss']), rtol=1e-5)


def test_csr_and_first_appearance():
    rng = np.random.default_rng(0)
    src, dst = rng.integers(0, 30, 200), rng.integers(0, 17, 200)
    indptr, indices, eperm = O.csr_by_dst(src, dst, 17)
    assert indptr[-1] == 200 and np.all(np.diff(indptr) == np.bincount(dst, minlength=17))
    for v in range(17):
        seg = eperm[indptr[v]:indptr[v + 1]]
        assert np.all(np.diff(seg) > 0) and np.all(dst[seg] == v) and np.all(src[seg] == indices[indptr[v]:indptr[v + 1]])
    ids, uniq = O.first_appearance_ids(['b', 'a', 'b', 'c', 'a'])
    assert ids.tolist() == [0, 1, 0, 2, 1] and uniq == ['b', 'a', 'c']


@pytest.mark.parametrize('name', ['tiny_mean', 'small_mean_128'])
def test_popularity_recs_match_reference(name):
    """use_popularity branch (src/metrics.py:69-72): fixture written by the reference's own get_recs."""
    meta, z = load_case(name)
    pmeta, zp = load_case(name + '_pop')
    hu, hi = torch.from_numpy(z['emb/user']), torch.from_numpy(z['emb/item'])
    buys = case_relations(z)[('user', 'buys', 'item')]
    bought = O.create_already_bought(buys[0], buys[1])
    uids = z['user_ids'].tolist()
    recs = O.get_recs(hu, hi, meta['k'], uids, bought, popularity=zp['popularity'], weight_popularity=pmeta['weight'])
    got = np.stack([np.asarray(recs[u], dtype=np.int64) for u in uids])
    cos = O.get_recs_scores(hu, hi, uids).numpy()
    ratings = np.stack([O.softmax(r) for r in cos]) + zp['popularity'].reshape(1, -1) * pmeta['weight']
    assert_topk_equivalent(got, zp['recs_pop'], ratings, meta['k'], tol=1e-7)


@pytest.mark.parametrize('name', ['tiny_mean', 'small_mean_128'])
def test_metrics_match_reference(name):
    """recs_to_metrics / get_metrics_at_k against fixtures written by the reference's own src/metrics.py:81-134
    (make_golden.py metrics_case): k of the base case and a large k (13 of 20 items; 40 > 32), with and without the
    already-bought filter, plus ragged hand-made recommendation lists with an empty one."""
    meta, z = load_case(name)
    mmeta, zm = load_case(name + '_metrics')
    hu, hi = torch.from_numpy(z['emb/user']), torch.from_numpy(z['emb/item'])
    g_users, g_items = zm['gt_users'], zm['gt_items']
    truth = O.create_already_bought(g_users, g_items)          # same dict-of-lists shape as create_ground_truth
    uids = np.unique(g_users).tolist()
    buys = case_relations(z)[('user', 'buys', 'item')]
    eids = zm['bought_eids']
    bought = O.create_already_bought(buys[0][eids], buys[1][eids])
    for kk in (mmeta['k'], mmeta['k_big']):
        for rm in (True, False):
            recs = O.get_recs(hu, hi, kk, uids, bought, remove_already_bought=rm)
            got = O.recs_to_metrics({u: [int(i) for i in v] for u, v in recs.items()}, truth, meta['n_items'])
            np.testing.assert_allclose(got, zm['metrics/k%d/remove%d' % (kk, int(rm))], rtol=0, atol=1e-12)
    lens, flat = zm['ragged/lens'], zm['ragged/items']
    off = np.concatenate([[0], np.cumsum(lens)])
    ragged = {int(u): flat[off[r]:off[r + 1]].tolist() for r, u in enumerate(zm['ragged/users'].tolist())}
    np.testing.assert_allclose(O.recs_to_metrics(ragged, truth, meta['n_items']), zm['ragged/metrics'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('name', ['fwd_fanout_mean', 'fwd_fanout_mean_128'])
def test_max_margin_loss_mask_and_recency_match_reference(name):
    """remove_false_negative / use_recency branches of max_margin_loss (src/model.py:516-531) against losses computed by
    the reference's own function (make_golden.py loss_case); recency exists for 'buys' only (the KeyError branch)."""
    meta, z = load_case(name)
    lmeta, zl = load_case(name + '_loss')
    pos, neg, mask = {}, {}, {}
    for c in RELS:
        if 'mask/%s' % c[1] in zl.files:
            pos[c], neg[c] = torch.from_numpy(z['pos/%s/score' % c[1]]), torch.from_numpy(z['neg/%s/score' % c[1]])
            mask[c] = torch.from_numpy(zl['mask/%s' % c[1]])
    rec = {('user', 'buys', 'item'): torch.from_numpy(zl['recency/buys'])}
    for rfn in (False, True):
        for ur in (False, True):
            got = O.max_margin_loss(pos, neg, lmeta['delta'], lmeta['neg_k'], use_recency=ur, recency_scores=rec,
                                    remove_false_negative=rfn, negative_mask=mask)
            np.testing.assert_allclose(float(got), float(zl['loss/mask%d/recency%d' % (int(rfn), int(ur))]), rtol=1e-6)
    assert float(zl['loss/mask1/recency1']) != float(zl['loss/mask0/recency0'])


def test_recs_to_metrics_formula():
    recs = {0: [1, 2, 3], 1: [4, 5, 6]}
    truth = {0: [2, 2, 9], 1: [7]}
    p, r, c = O.recs_to_metrics(recs, truth, 10)
    assert p == 1 / 6 and r == 2 / 4 and c == 0.6


@pytest.mark.parametrize('name', ['sport_mean_edge', 'sport_pool_nn'])
def test_three_node_type_schema_matches_reference(name):
    """10-relation schema with 'sport' nodes (src/utils_data.py:204-238); only users and items are seeded, so the
    sport table stays zero (run.py:329-333); `_edge` weighting applies to user-item relations only (model.py:173)."""
    meta, z = load_case(name)
    num = meta['num']
    rels = [tuple(c) for c in meta['rels']]
    coo = {c: (z['edges/%s/src' % c[1]], z['edges/%s/dst' % c[1]],
               z['occurrence/%s' % c[1]] if 'occurrence/%s' % c[1] in z.files else None) for c in rels}
    blk = O.block_from_coo(num, num, coo)
    feats = {t: torch.from_numpy(z['feat/' + t]) for t in num}
    seeds = {'user': np.arange(num['user']), 'item': np.arange(num['item'])}
    y = O.get_embeddings_full(num, [blk] * (meta['n_layers'] - 1), feats, state_dict(z), meta['out'], seeds,
                              meta['aggregator'])
    for t in num:
        np.testing.assert_allclose(y[t].numpy(), z['emb/' + t], rtol=1e-4, atol=1e-5)
    assert float(np.abs(z['emb/sport']).max()) == 0.0
