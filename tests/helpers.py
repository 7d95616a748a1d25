"""Shared test helpers: golden-fixture loading and tie-aware top-k comparison."""
import json
import os

import numpy as np
import torch

from oracle import straightline as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
RELS = [('item', 'bought-by', 'user'), ('item', 'clicked-by', 'user'), ('user', 'buys', 'item'), ('user', 'clicks', 'item')]
EMBED_CASES = ['tiny_mean', 'tiny_mean_nn', 'tiny_pool_nn', 'tiny_mean_edge', 'tiny_pool_nn_edge', 'tiny_mean_nonorm',
               'tiny_mean_noembed', 'tiny_pool_hetero_max', 'tiny_mean_hetero_mean', 'tiny_mean_batched',
               'small_mean_128', 'small_pool_256', 'preset_64_32', 'preset_192_96', 'preset_512_256']


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    return meta, z


def state_dict(z):
    return {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}


def case_relations(z):
    rel = {}
    for c in RELS:
        rel[c] = (z['edges/%s/src' % c[1]], z['edges/%s/dst' % c[1]])
    return rel


def case_occurrence(z):
    return {c: z['occurrence/%s' % c[1]] for c in RELS if 'occurrence/%s' % c[1] in z.files}


def assert_topk_equivalent(got, want, scores, k, tol=1e-5):
    """``got`` / ``want``: [n, k] item ids (-1 = empty). Rows must agree position by position except where the
    fp32 scores of the two ids differ by less than ``tol`` (ties are unordered in the reference: np.argsort)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    for r in range(got.shape[0]):
        if np.array_equal(got[r], want[r]):
            continue
        assert np.array_equal(got[r] < 0, want[r] < 0), 'row %d: different number of recommendations' % r
        valid = want[r] >= 0
        sg, sw = scores[r, got[r][valid]], scores[r, want[r][valid]]
        assert np.all(np.abs(sg - sw) < tol), 'row %d: ids differ beyond score ties: %s vs %s' % (r, got[r], want[r])
        assert len(set(got[r][valid].tolist())) == int(valid.sum()), 'row %d: duplicate ids' % r


def oracle_blocks(g, sampler, seeds, key, exclude=None):
    """Blocks restated by the oracle (sample_frontier + compact_block), innermost layer first."""
    cets = g.canonical_etypes
    seeds = {t: np.asarray(v, np.int64) for t, v in seeds.items() if len(v)}
    out = []
    for layer in reversed(range(sampler.num_layers)):
        fr = {}
        for ci, c in enumerate(cets):
            if c[2] in seeds and g.num_edges(c):
                indptr, indices, eperm = g.csr(c)
                fr[c] = O.sample_frontier(indptr, indices, eperm, seeds[c[2]], sampler._fanout(layer),
                                          O.sample_key(key, layer * 64 + ci), (exclude or {}).get(c, ()))
        src_ids, rels = O.compact_block(g.ntypes, cets, seeds, fr)
        out.insert(0, (src_ids, rels, dict(seeds)))
        seeds = {t: v for t, v in src_ids.items() if v.size}
    return out


def assert_blocks_equal_oracle(blocks, want, g):
    for b, (src_ids, rels, seeds) in zip(blocks, want):
        for t in g.ntypes:
            assert b.srcnodes[t].data['_ID'].cpu().tolist() == src_ids[t].tolist()
            assert b.dstnodes[t].data['_ID'].cpu().tolist() == seeds.get(t, np.zeros(0)).tolist()
            assert b.num_src[t] == src_ids[t].size and b.num_dst[t] == (seeds[t].size if t in seeds else 0)
            for name, v in g.nodes[t].data.items():
                assert torch.equal(b.srcnodes[t].data[name].cpu(), v[torch.from_numpy(src_ids[t])])
        for c in g.canonical_etypes:
            r = b.rels[c]
            if c in rels:
                ip, ls, e = rels[c]
                assert r.indptr.cpu().tolist() == ip.tolist() and r.indices.cpu().tolist() == ls.tolist()
                assert r.eperm.cpu().tolist() == e.tolist()
            else:
                assert r.nnz == 0 and r.indptr.cpu().tolist() == [0] * (r.n_dst + 1)
            assert r.n_src == b.num_src[c[0]] and r.n_dst == b.num_dst[c[2]]
