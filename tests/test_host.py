"""Host-side logic (no GPU): graph containers, block construction, bought-list CSR, error-bound arithmetic."""
import numpy as np
import pytest
import torch

import gnn_recsys_b200 as grb
from oracle import straightline as O
from helpers import oracle_blocks, assert_blocks_equal_oracle


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


def test_loaders_carry_occurrence_reject_unknown_kwargs_and_stay_lazy():
    """Drop-in details of the loaders: (1) blocks carry the parent graph's edata['occurrence'] by default, like DGL's
    (the `*_edge` aggregators read it, src/model.py:174), on the NodeDataLoader and the EdgeDataLoader; (2) unknown
    keyword arguments raise instead of being swallowed, torch DataLoader's own are accepted; (3) a full-graph
    NodeDataLoader hands out a lazy block: iterating it builds no host CSR."""
    d, g = small_graph()
    rng = np.random.default_rng(3)
    occ = {}
    for et, twin in (('buys', 'bought-by'), ('clicks', 'clicked-by')):
        occ[et] = occ[twin] = torch.from_numpy(rng.integers(1, 5, g.num_edges(et)))
    for et, v in occ.items():
        g.edges[et].data['occurrence'] = v
    eids = {'buys': np.arange(g.num_edges('buys')), 'clicks': np.arange(g.num_edges('clicks'))}
    rev = {'buys': 'bought-by', 'bought-by': 'buys', 'clicks': 'clicked-by', 'clicked-by': 'clicks'}
    loader = grb.EdgeDataLoader(g, eids, grb.MultiLayerNeighborSampler([5, 5]), exclude='reverse_types', reverse_etypes=rev,
                                negative_sampler=grb.negative_sampler.Uniform(3), batch_size=32, shuffle=True,
                                num_workers=0, pin_memory=False)
    assert loader.edge_weight == 'occurrence'
    _, pos_g, neg_g, blocks = next(iter(loader))
    for b in blocks:
        for c, r in b.rels.items():
            assert r.weight is not None and r.weight.dtype == torch.float32 and r.weight.shape[0] == r.nnz
            assert torch.equal(r.weight, occ[c[1]][r.eperm.long()].float())   # the weight of the edge held by each CSR slot
    off = grb.EdgeDataLoader(g, eids, grb.MultiLayerNeighborSampler([5, 5]), batch_size=32, edge_weight=None)
    assert all(r.weight is None for b in next(iter(off))[3] for r in b.rels.values())
    nl = grb.NodeDataLoader(g, {'user': np.arange(60)}, grb.MultiLayerNeighborSampler([4]), batch_size=16)
    assert all(r.weight is not None for r in next(iter(nl))[2][0].rels.values())
    with pytest.raises(TypeError):
        grb.EdgeDataLoader(g, eids, grb.MultiLayerNeighborSampler([5]), batch_size=8, edge_wieght='occurrence')
    with pytest.raises(TypeError):
        grb.NodeDataLoader(g, {'user': np.arange(60)}, grb.MultiLayerFullNeighborSampler(1), batchsize=8)
    _, g2 = small_graph(seed=1)
    assert grb.NodeDataLoader(g2, {'user': np.arange(60)}, grb.MultiLayerFullNeighborSampler(1)).edge_weight is None
    full = grb.NodeDataLoader(g2, {'user': np.arange(60), 'item': np.arange(25)}, grb.MultiLayerFullNeighborSampler(2),
                              batch_size=128, shuffle=True)
    assert full.full_graph
    (_, out_nodes, blks), = list(full)
    assert len(blks) == 2 and not g2._csr_cache                 # nothing built on the host ...
    assert blks[0].number_of_dst_nodes('user') == 60 and g2._csr_cache   # ... until somebody actually reads the block
    with pytest.raises(ValueError):
        grb.HeteroGraph({('user', 'buys', 'item'): (np.array([0, -1]), np.array([1, 1]))})


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
    """recs.score_err_bound (the host mirror of the device formula) really bounds the error of each product scheme,
    emulated on the CPU with the measured residuals of the rounded rows."""
    g = torch.Generator().manual_seed(0)
    x = torch.nn.functional.normalize(torch.rand(256, 128, generator=g), dim=1)
    y = torch.nn.functional.normalize(torch.rand(512, 128, generator=g), dim=1)
    y = y - y.mean(0)
    exact = x.double() @ y.double().t()
    for elem, dt in (('bf16', torch.bfloat16), ('fp16', torch.float16)):
        def split(v):
            hi = v.to(dt).float()
            return hi, (v - hi).to(dt).float()
        xh, xl = split(x)
        yh, yl = split(y)
        d = lambda a, b: a.double() @ b.double().t()
        for pu, pi, approx in ((1, 1, d(xh, yh)), (2, 1, d(xh, yh) + d(xl, yh)), (2, 2, d(xh, yh) + d(xl, yh) + d(xh, yl))):
            xr1, yr1 = (x - xh).norm(dim=1), (y - yh).norm(dim=1)
            xr = xr1 if pu == 1 else (x - xh - xl).norm(dim=1)
            yr = yr1 if pi == 1 else (y - yh - yl).norm(dim=1)
            stats = [float(y.norm(dim=1).max()), 1.0, float(yr.max()), float(yr1.max())]
            bound = grb.recs.score_err_bound(float(xr.max()), float(xr1.max()), stats, elem, pu, pi, 0.0)
            err = (approx - exact).abs().max().item()
            assert err <= bound, (elem, pu, pi, err, bound)
            if (pu, pi) == (2, 2):
                assert bound < (2e-5 if elem == 'bf16' else 1e-6)
    c = grb.RecsConfig()
    assert (c.elem, c.products, c.shortlist, c.small_items) == ('fp16', 1, 32, 32768) and c.second == ('fp16', 2, 2, 16)
    assert grb.RecsConfig(parts=2, elem='bf16').products == 3 and grb.RecsConfig(parts=2).second is None
    with pytest.raises(ValueError):
        grb.RecsConfig(parts_users=1, parts_items=2)


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


# ------------------------------------------------------------------------------------------------ sampled blocks
def test_counter_based_hash_matches_oracle_and_library_key():
    import ctypes as C
    from gnn_recsys_b200 import _native as N
    lib = C.CDLL(N.LIB_PATH)  # host-only entry point: no GPU needed
    lib.gr_sample_key.restype, lib.gr_sample_key.argtypes = C.c_uint64, [C.c_uint64, C.c_uint64]
    rng = np.random.default_rng(0)
    for seed, stream in [(0, 0), (1, 5), (2 ** 63 - 1, 4099), (int(rng.integers(0, 2 ** 63)), 130)]:
        k = grb.sample_key(seed, stream)
        assert k == O.sample_key(seed, stream) == lib.gr_sample_key(seed, stream)
        ctr = np.concatenate([np.arange(50), rng.integers(0, 2 ** 62, 50)]).astype(np.int64)
        assert grb.hash64(k, ctr).tolist() == [O.hash64(k, int(c)) for c in ctr]


@pytest.mark.parametrize('fanouts', [[3, 2], [1], None])
def test_host_sampler_matchesoracle_blocks(fanouts):
    d, g = small_graph(5, 900)
    sampler = grb.MultiLayerNeighborSampler(fanouts) if fanouts else grb.MultiLayerFullNeighborSampler(2)
    seeds = {'user': np.array([7, 3, 59, 0]), 'item': np.array([24, 1, 2])}
    excl = {('user', 'buys', 'item'): np.array([0, 5, 9, 11]), ('item', 'bought-by', 'user'): np.array([0, 5, 9, 11])}
    for key, ex in ((12345, None), (2 ** 62 + 17, excl)):
        blocks = sampler.sample_blocks(g, seeds, key=key, exclude=ex)
        assert_blocks_equal_oracle(blocks, oracle_blocks(g, sampler, seeds, key, ex), g)
        for b in blocks:  # excluded edges never show up, fan-out bound holds
            for c, r in b.rels.items():
                if ex and c in ex:
                    assert not set(r.eperm.tolist()) & set(ex[c].tolist())
                if fanouts:
                    assert int(np.diff(r.indptr.numpy()).max(initial=0)) <= max(fanouts)


def test_fanout_sampling_is_uniform_without_replacement():
    """Every in-edge of a row is kept with probability fanout / degree (over keys); never twice in one draw."""
    src = np.arange(40) % 13
    g = grb.HeteroGraph({('user', 'buys', 'item'): (src, np.zeros(40, np.int64))}, {'user': 13, 'item': 1})
    sampler = grb.MultiLayerNeighborSampler([8])
    hits = np.zeros(40)
    n = 600
    for key in range(n):
        r = sampler.sample_blocks(g, {'item': np.array([0])}, key=key)[0].rels[('user', 'buys', 'item')]
        e = r.eperm.numpy()
        assert e.size == 8 and np.unique(e).size == 8 and np.all(np.diff(e) > 0)
        hits[e] += 1
    p = hits / n
    assert abs(p.mean() - 0.2) < 1e-9 and np.all(np.abs(p - 0.2) < 5 * np.sqrt(0.2 * 0.8 / n))


def test_negative_sampler_matches_oracle_and_is_uniform():
    d, g = small_graph(6, 500)
    c = ('user', 'clicks', 'item')
    eids = np.array([3, 0, 17, 3])
    got = grb.negative_sampler.Uniform(7)(g, {'clicks': eids}, 99)[c]
    want = O.negative_uniform(g.edge_arrays(c)[0], eids, 7, g.num_nodes('item'), O.sample_key(99, 4096 + g.canonical_etypes.index(c)))
    assert got[0].tolist() == want[0].tolist() and got[1].tolist() == want[1].tolist()
    big = grb.negative_sampler.Uniform(2000)(g, {'clicks': np.arange(50)}, 5)[c][1]
    counts = np.bincount(big, minlength=25)
    assert counts.sum() == 100000 and np.all(np.abs(counts - 4000) < 5 * np.sqrt(4000))


def test_device_block_builder_host_logic_with_oracle_kernels(monkeypatch):
    """The Python orchestration of ``sampling_device.py`` (buffer offsets, remap splitting, block assembly) with the
    four C-ABI calls it makes replaced by the ORACLE's restatements, on CPU tensors: must equal the host builder.
    (The kernels themselves are checked against the same oracle functions on the GPU, tests/test_gpu_parity.py.)"""
    import importlib
    sd_mod = importlib.import_module('gnn_recsys_b200.sampling_device')
    ops = grb.ops
    cpu = torch.device('cpu')

    def fake_count(indptr, eperm, seeds, fanout, excl=None):
        ip, _, _ = O.sample_frontier(indptr.numpy(), np.zeros(int(indptr[-1]), np.int64),
                                     None if eperm is None else eperm.numpy(), seeds.numpy(), fanout, 0,
                                     () if excl is None else excl.numpy())
        return torch.from_numpy(ip), torch.tensor([int(ip[-1])], dtype=torch.int32)

    def fake_fill(indptr, indices, eperm, seeds, fanout, excl, key, out_indptr, out_src, out_eid):
        ip, s, e = O.sample_frontier(indptr.numpy(), indices.numpy(), None if eperm is None else eperm.numpy(),
                                     seeds.numpy(), fanout, key, () if excl is None else excl.numpy())
        assert ip.tolist() == out_indptr.tolist()
        out_src.copy_(torch.from_numpy(s))
        out_eid.copy_(torch.from_numpy(e))

    def fake_neg(edge_src, eids, k, n_dst_nodes, key):
        s, d = O.negative_uniform(edge_src.numpy(), eids.numpy(), k, n_dst_nodes, key)
        return torch.from_numpy(s), torch.from_numpy(d)

    def fake_remap(raw):
        ids, uniq = O.first_appearance_ids(raw.tolist())
        return torch.from_numpy(ids.astype(np.int32)), torch.tensor(uniq, dtype=torch.int64)

    def fake_full_block_on(self, device, edge_weight=None):
        return self.full_block(edge_weight, with_features=False)

    def fake_device_edges(self, etype, device):
        s, d = self.edge_arrays(etype)
        return torch.from_numpy(s.astype(np.int32)), torch.from_numpy(d.astype(np.int32))

    monkeypatch.setattr(ops, 'sample_count', fake_count)
    monkeypatch.setattr(ops, 'sample_fill', fake_fill)
    monkeypatch.setattr(ops, 'negative_uniform', fake_neg)
    monkeypatch.setattr(ops, 'remap_first_appearance', fake_remap)
    monkeypatch.setattr(ops, 'remap_many', lambda raws: [fake_remap(r) for r in raws])
    monkeypatch.setattr(grb.HeteroGraph, 'full_block_on', fake_full_block_on)
    monkeypatch.setattr(grb.HeteroGraph, 'device_edges', fake_device_edges)

    d, g = small_graph(8, 1500)
    rng = np.random.default_rng(0)
    for fwd, bwd in (('buys', 'bought-by'), ('clicks', 'clicked-by')):
        occ = torch.from_numpy(rng.integers(1, 5, g.num_edges(fwd)).astype(np.float32))
        g.edges[fwd].data['occurrence'], g.edges[bwd].data['occurrence'] = occ, occ
    rev = {'buys': 'bought-by', 'bought-by': 'buys', 'clicks': 'clicked-by', 'clicked-by': 'clicks'}
    eids = {'buys': np.arange(g.num_edges('buys')), 'clicks': np.arange(g.num_edges('clicks'))}
    for sampler in (grb.MultiLayerNeighborSampler([4, 3]), grb.MultiLayerFullNeighborSampler(2)):
        kw = dict(exclude='reverse_types', reverse_etypes=rev, negative_sampler=grb.negative_sampler.Uniform(6),
                  batch_size=40, shuffle=True, seed=3)
        host = grb.EdgeDataLoader(g, eids, sampler, **kw)
        devl = grb.EdgeDataLoader(g, eids, sampler, **kw)
        it_h = iter(host)
        order = devl.rng.permutation(devl._flat_e.size)
        for b in range(2):
            in_h, pos_h, neg_h, blocks_h = next(it_h)
            # drive the device builder directly (the loader's own dispatch insists on a CUDA device)
            sel = order[b * 40:(b + 1) * 40]
            items = {c: devl._flat_e[sel[devl._flat_t[sel] == i]] for i, c in enumerate(devl._types)}
            items = {c: e for c, e in items.items() if e.size}
            key = int(devl.rng.integers(0, 2 ** 63, dtype=np.int64))
            in_d, pos_d, neg_d, blocks_d = sd_mod.edge_batch_device(devl, items, key, cpu)
            for c in g.canonical_etypes:
                for a, b in ((pos_h, pos_d), (neg_h, neg_d)):
                    for x, y in zip(a.edge_arrays(c), b.edge_arrays(c)):
                        assert x.tolist() == y.tolist()
            for bh, bd in zip(blocks_h, blocks_d):
                assert bh.num_src == bd.num_src and bh.num_dst == bd.num_dst
                for c in g.canonical_etypes:
                    rh, rd = bh.rels[c], bd.rels[c]
                    assert rh.indptr.tolist() == rd.indptr.tolist() and rh.indices.tolist() == rd.indices.tolist()
                    assert rh.eperm.tolist() == rd.eperm.tolist() and (rh.n_src, rh.n_dst) == (rd.n_src, rd.n_dst)
                for t in g.ntypes:
                    assert torch.equal(bh.srcnodes[t].data['features'], bd.srcnodes[t].data['features'])
                    assert bh.srcnodes[t].data[grb.NID].tolist() == bd.srcnodes[t].data[grb.NID].tolist()
                    assert bh.dstnodes[t].data[grb.NID].tolist() == bd.dstnodes[t].data[grb.NID].tolist()
    # edge weights gathered by edge id (the *_edge aggregators): node-loader blocks, host vs device builder
    seeds = {'user': np.array([5, 1, 40]), 'item': np.array([3, 2])}
    s = grb.MultiLayerNeighborSampler([5, 5])
    bh = s.sample_blocks(g, seeds, key=77, edge_weight='occurrence')
    bd = sd_mod.sample_blocks_device(g, s, seeds, 77, cpu, None, 'occurrence')
    for x, y in zip(bh, bd):
        for c in g.canonical_etypes:
            assert x.rels[c].indices.tolist() == y.rels[c].indices.tolist()
            assert torch.equal(x.rels[c].weight, y.rels[c].weight)


def test_bench_tensor_peak_follows_the_measured_clock():
    """bench.py reports the scoring kernel against the cuBLAS figure of the clock regime the run was in: burst when the
    median SM clock under load is nearer the maximum than the clock the sustained figure was measured at."""
    import bench
    pk = dict(hbm=6552.0, tc_burst=1661.5, tc=1389.2, source='measured', sustained_mhz=1350.0, max_mhz=1965.0)
    assert bench.tensor_peak(pk, {'sm_mhz': 1875.0, 'sm_max_mhz': 1965.0})[0] == 1661.5
    peak, label = bench.tensor_peak(pk, {'sm_mhz': 1530.0, 'sm_max_mhz': 1965.0})
    assert peak == 1389.2 and 'sustained' in label
    assert bench.tensor_peak(pk, None)[0] == 1389.2                       # no clock samples: the conservative figure
    fb = dict(hbm=6650.0, tc_burst=1590.0, tc=1400.0, source='fallback', sustained_mhz=None, max_mhz=None)
    assert bench.tensor_peak(fb, {'sm_mhz': 1900.0, 'sm_max_mhz': 1965.0})[0] == 1590.0
    assert bench.tensor_peak(fb, {'sm_mhz': 1300.0, 'sm_max_mhz': 1965.0})[0] == 1400.0


def test_ncu_traffic_tool_reads_wide_and_long_csv(tmp_path):
    """tools/ncu_traffic.py: per-stage DRAM bytes from an `ncu --page raw --csv` (wide) or `ncu --metrics ... --csv`
    (long) capture; a partial capture makes bench.dram_traffic() return None instead of an undercount."""
    import json, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    wide = tmp_path / 'wide.csv'
    wide.write_text('"ID","Kernel Name","dram__bytes_read.sum","dram__bytes_write.sum","gpu__time_duration.sum"\n'
                    '"","","byte","byte","ns"\n'
                    '"0","void <unnamed>::sage_fused_kernel<1, 1, 4, 2, 0, 1>(int)","1000","500","2000000"\n'
                    '"1","void <unnamed>::score_topk_kernel<2, 1, 1, 1>(int)","300","100","30000000"\n'
                    '"2","<unnamed>::order_cosine_kernel(int)","40","2","1000"\n')
    long_ = tmp_path / 'long.csv'
    rows = ['"ID","Kernel Name","Metric Name","Metric Unit","Metric Value"']
    for i, (name, rd, wr, ns) in enumerate([('void <unnamed>::long_partial_kernel<1, 0>(int)', 7000, 10, 4000000),
                                            ('<unnamed>::rescore_kernel(int)', 50, 5, 2000000)]):
        for m, u, v in (('dram__bytes_read.sum', 'byte', rd), ('dram__bytes_write.sum', 'byte', wr),
                        ('gpu__time_duration.sum', 'ns', ns)):
            rows.append('"%d","%s","%s","%s","%s"' % (i, name, m, u, format(v, ',')))
    long_.write_text('==PROF== noise line\n' + '\n'.join(rows) + '\n')
    out = tmp_path / 'traffic.json'
    tool = os.path.join(root, 'tools', 'ncu_traffic.py')
    subprocess.run([sys.executable, tool, str(wide), 'cA', '--out', str(out)], check=True, capture_output=True)
    subprocess.run([sys.executable, tool, str(long_), 'cB', '--partial', '--out', str(out)], check=True, capture_output=True)
    t = json.loads(out.read_text())
    assert t['cA']['aggregate']['dram_bytes'] == 1500 and t['cA']['score']['dram_bytes'] == 400
    assert t['cA']['prep']['dram_bytes'] == 42 and abs(t['cA']['score']['kernel_ms'] - 30.0) < 1e-9
    assert t['cB']['aggregate']['dram_bytes'] == 7010 and t['cB']['rescore']['launches'] == 1
    assert t['cB']['aggregate']['partial'] is True and t['cA']['aggregate']['partial'] is False
