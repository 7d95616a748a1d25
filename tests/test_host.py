"""Host-side logic (no GPU): graph containers, block construction, bought-list CSR, error-bound arithmetic."""
import numpy as np
import pytest
import torch

import gnn_recsys_b200 as grb
from oracle import straightline as O


def small_graph(seed=0, n_edges=400):
    d = grb.make_graph(60, 25, n_edges, seed)
    return d, d.graph()


def test_heterograph_surface_matches_dgl_rules():
    d, g = small_graph()
    assert g.ntypes == ['item', 'user']
    assert g.canonical_etypes == sorted(g.canonical_etypes) and len(g.canonical_etypes) == 4
    assert g.num_nodes('user') == 60 and g.num_nodes('item') == 25
    nb = int(d.is_buy.sum())
    assert g.num_edges('buys') == nb and g.num_edges('bought-by') == nb
    u, v = g.find_edges(torch.tensor([0, 3]), etype='buys')
    s, t = d.relations()[('user', 'buys', 'item')]
    assert u.tolist() == s[[0, 3]].tolist() and v.tolist() == t[[0, 3]].tolist()
    eids = g.out_edges(torch.tensor([1, 2]), form='eid', etype='buys')
    assert set(s[eids.numpy()].tolist()) <= {1, 2}
    with pytest.raises(KeyError):
        g.to_canonical_etype('likes')
    with pytest.raises(ValueError):
        grb.HeteroGraph({('user', 'buys', 'item'): (np.array([5]), np.array([1]))}, {'user': 3, 'item': 2})


def test_host_csr_bit_exact_against_oracle():
    rng = np.random.default_rng(1)
    src, dst = rng.integers(0, 100, 3000), rng.integers(0, 37, 3000)
    a = grb.csr_by_dst_host(src, dst, 37)
    b = O.csr_by_dst(src, dst, 37)
    for x, y in zip(a, b):
        assert x.dtype == np.int32 and np.array_equal(x, y)
    e = grb.csr_by_dst_host(np.zeros(0, np.int64), np.zeros(0, np.int64), 5)
    assert e[0].tolist() == [0] * 6 and e[1].size == 0
    with pytest.raises(IndexError):
        grb.csr_by_dst_host(np.array([0]), np.array([9]), 5)


def test_full_block_and_minibatch_blocks_agree_with_oracle_embeddings():
    """Blocks built by the host samplers feed the ORACLE to the same embeddings as the full-graph block."""
    d, g = small_graph(3, 3000)  # dense enough that every mini-batch has edges in all four relations (SURVEY 8a, a6 hazard)
    torch.manual_seed(0)
    dims = {'user': 2, 'item': 4, 'hidden': 8, 'out': 8}
    sd = {}
    for t in ('user', 'item'):
        sd['%s_embed.proj_feats.weight' % t] = torch.randn(8, dims[t])
        sd['%s_embed.proj_feats.bias' % t] = torch.randn(8)
    for li in range(2):
        for et in ('buys', 'bought-by', 'clicks', 'clicked-by'):
            for nm in ('fc_self', 'fc_neigh'):
                sd['layers.%d.mods.%s.%s.weight' % (li, et, nm)] = torch.randn(8, 8) * 0.4
    num = {'user': 60, 'item': 25}
    blk = O.block_from_coo(num, num, {c: (s.astype(np.int64), t.astype(np.int64), None) for c, (s, t) in d.relations().items()})
    feats = {'user': d.user_feat, 'item': d.item_feat}
    want = O.get_embeddings_full(num, [blk, blk], feats, sd, 8)
    loader = grb.NodeDataLoader(g, {'user': np.arange(60), 'item': np.arange(25)}, grb.MultiLayerFullNeighborSampler(2),
                                batch_size=16, shuffle=True, seed=1, force_minibatch=True)
    assert len(loader) == 6 and not loader.full_graph
    got = {t: torch.zeros(n, 8) for t, n in num.items()}
    for _, out_nodes, blocks in loader:
        obs = []
        for b in blocks:
            rels = {}
            for c, r in b.rels.items():
                dst = np.repeat(np.arange(r.n_dst), np.diff(r.indptr.numpy()))
                rels[c] = (r.indices.numpy().astype(np.int64), dst, None)
            obs.append(O.block_from_coo(b.num_src, b.num_dst, rels))
            for t in b.dsttypes:  # DGL block invariant: destination nodes are the first source nodes
                nd = b.num_dst[t]
                assert b.srcnodes[t].data[grb.NID][:nd].tolist() == b.dstnodes[t].data[grb.NID].tolist()
        h = O.get_repr(obs, O.embed_inputs({t: v for t, v in blocks[0].srcdata['features'].items()}, sd), sd)
        for t in h.keys():  # like src/train/run.py:347-348
            got[t][out_nodes[t]] = h[t]
    for t in num:
        np.testing.assert_allclose(got[t].numpy(), want[t].numpy(), rtol=1e-4, atol=1e-5)
    full = grb.NodeDataLoader(g, {'user': np.arange(60), 'item': np.arange(25)}, grb.MultiLayerFullNeighborSampler(2))
    assert full.full_graph and len(full) == 1


def test_edge_loader_layout():
    d, g = small_graph(4)
    eids = {'buys': np.arange(g.num_edges('buys')), 'clicks': np.arange(g.num_edges('clicks'))}
    rev = {'buys': 'bought-by', 'bought-by': 'buys', 'clicks': 'clicked-by', 'clicked-by': 'clicks'}
    loader = grb.EdgeDataLoader(g, eids, grb.MultiLayerNeighborSampler([3, 3]), exclude='reverse_types', reverse_etypes=rev,
                                negative_sampler=grb.negative_sampler.Uniform(5), batch_size=32, shuffle=True, seed=2)
    _, pos_g, neg_g, blocks = next(iter(loader))
    assert len(blocks) == 2
    npos = sum(pos_g.num_edges(c) for c in pos_g.canonical_etypes)
    nneg = sum(neg_g.num_edges(c) for c in neg_g.canonical_etypes)
    assert npos == 32 and nneg == 32 * 5
    for c in (('user', 'buys', 'item'), ('user', 'clicks', 'item')):  # K-consecutive negatives share the positive's source
        ps, _ = pos_g.edge_arrays(c)
        ns, _ = neg_g.edge_arrays(c)
        assert np.array_equal(np.repeat(ps, 5), ns)
    for t in g.ntypes:
        assert blocks[-1].num_dst[t] == pos_g.num_nodes(t)
    for b in blocks:  # fan-out bound
        for r in b.rels.values():
            assert int(np.diff(r.indptr.numpy()).max(initial=0)) <= 3


def test_bought_csr_forms_agree():
    users = np.array([3, 1, 3, 0, 3, 1])
    items = np.array([9, 4, 2, 7, 9, 4])
    csr = grb.BoughtCSR.from_edges(users, items, 5)
    assert csr[3] == [2, 9, 9] and csr[1] == [4, 4] and csr[2] == [] and csr.n_rows == 5
    d = O.create_already_bought(users, items)
    sub = grb.BoughtCSR.from_dict(d, [3, 2, 0])
    assert sub.n_rows == 3 and sub[3] == [2, 9, 9] and sub[2] == [] and sub[0] == [7]
    sel = csr.select([3, 2, 0])
    assert np.array_equal(sel.indptr, sub.indptr) and np.array_equal(sel.ids, sub.ids)
    plain = grb.BoughtCSR.from_dict({1: [5]}, [0, 1])
    assert plain[0] == [] and plain[1] == [5]
    assert csr.select(np.arange(5)) is csr


def test_error_bound_arithmetic():
    c = grb.RecsConfig(elem='bf16', parts=2, acc_err=0.0)
    assert abs(c.err_rel() - 3 * 2.0 ** -18) < 1e-7
    assert grb.RecsConfig(elem='fp16', parts=2, acc_err=0.0).err_rel() < 8e-7
    assert grb.RecsConfig(elem='bf16', parts=1, acc_err=0.0).err_rel() > 3.9e-3
    assert grb.RecsConfig(elem='fp16').err_abs(128) > 0 and grb.RecsConfig(elem='bf16').err_abs(128) == 0
    # empirical check of the split-product bound on the CPU (bf16 hi/lo emulation)
    g = torch.Generator().manual_seed(0)
    x = torch.nn.functional.normalize(torch.rand(256, 128, generator=g), dim=1)
    y = torch.nn.functional.normalize(torch.rand(512, 128, generator=g), dim=1)
    y = y - y.mean(0)

    def split(v):
        hi = v.to(torch.bfloat16).float()
        return hi, (v - hi).to(torch.bfloat16).float()
    xh, xl = split(x)
    yh, yl = split(y)
    approx = (xh.double() @ yh.double().t()) + (xl.double() @ yh.double().t()) + (xh.double() @ yl.double().t())
    err = (approx - x.double() @ y.double().t()).abs().max().item()
    assert err <= c.err_rel() * float(y.norm(dim=1).max())


def test_unsupported_options_fail_loudly():
    _, g = small_graph()
    dims = {'user': 2, 'item': 4, 'hidden': 8, 'out': 8}
    with pytest.raises(NotImplementedError):
        grb.ConvModel(g, 2, dims, aggregator_type='lstm')
    with pytest.raises(NotImplementedError):
        grb.ConvModel(g, 2, dims, pred='nn')
    with pytest.raises(KeyError):
        grb.ConvModel(g, 2, dims, pred='dot')
    with pytest.raises(KeyError):
        grb.ConvModel(g, 2, dims, aggregator_hetero='median')
    m = grb.ConvModel(g, 3, dims, aggregator_type='pool_nn')
    keys = set(m.state_dict().keys())
    assert 'user_embed.proj_feats.weight' in keys and 'item_embed.proj_feats.bias' in keys
    assert 'layers.0.mods.bought-by.fc_preagg.weight' in keys and 'layers.1.mods.clicks.fc_neigh.weight' in keys
    assert len(m.layers) == 2
    assert len(grb.ConvModel(g, 2, dims, embedding_layer=False).layers) == 2
