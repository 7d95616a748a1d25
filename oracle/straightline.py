"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GNN-RecSys hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this
module; the product package (``gnn-recsys_b200/``) never does and has no CPU fallback.

A straight-line torch/NumPy (CPU, fp32 or fp64) restatement of the reference algorithm, function by
function, each citing the reference lines it follows (paths relative to the reference repository root):

  a1  NodeEmbedding.forward            src/model.py:10-24
  a3-a5 ConvLayer.forward              src/model.py:123-237   (mean / mean_nn / pool_nn and their *_edge forms)
  a6  HeteroGraphConv (dgl 0.5.2)      constructed at src/model.py:384-406; semantics restated from DGL
  a8  ConvModel.get_repr               src/model.py:415-421
  a9  ConvModel.forward                src/model.py:423-470
  a10 CosinePrediction.forward         src/model.py:308-327
  a11 max_margin_loss                  src/model.py:473-533
  a12 get_embeddings                   src/train/run.py:311-349
  a13 get_recs                         src/metrics.py:31-78
  a14 create_already_bought            src/metrics.py:19-28

PARITY STATUS. The arithmetic of a3-a6 and a10 lives in ``dgl==0.5.2`` (requirements.txt:2), which is neither
vendored in the reference nor installable in this image, and the reference ships no tests or golden vectors:
for those rules parity is **unpinned** -- they restate DGL's published semantics (fn.mean = sum / clamp(deg, 1)
with zero rows for isolated nodes; fn.max with zero rows for isolated nodes; HeteroGraphConv skips empty
relations and reduces per destination type). Everything that is plain torch / NumPy in the reference is
**pinned**: ``tests/golden/make_golden.py`` imports the reference's own unmodified ``src/model.py``,
``src/train/run.py`` and ``src/metrics.py`` from ``/root/reference`` (on top of ``oracle/dgl_shim``) and the
committed fixtures under ``tests/golden/`` are its outputs; ``tests/test_oracle.py`` checks this file
against them.

Data conventions: a *block* is ``{'num_src': {nt: n}, 'num_dst': {nt: n}, 'rels': {(st, et, dt): (src, dst, w)}}``
with int64 COO ids local to the block and ``w`` an optional per-edge fp32 scalar; weights use the
reference's ``state_dict`` keys (``layers.{i}.mods.{etype}.fc_self.weight`` ...).
"""
from collections import defaultdict

import numpy as np
import torch

AGGREGATORS = ('mean', 'mean_nn', 'pool_nn', 'mean_edge', 'mean_nn_edge', 'pool_nn_edge')


# ----------------------------------------------------------------------------------------------- graph prep
def csr_by_dst(src, dst, n_dst):
    """Stable COO -> CSR over destination rows: neighbours of a row stay in edge-id order.
    (What DGL builds internally for ``update_all``; int32 layout of the product's ``gr_csr_build_i32``.)"""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    eid = np.arange(src.shape[0], dtype=np.int64)
    order = np.lexsort((eid, dst))  # primary key dst, ties by edge id
    indptr = np.zeros(n_dst + 1, dtype=np.int64)
    for v in dst:  # plain counting loop on purpose: independent of np.bincount / np.cumsum
        indptr[v + 1] += 1
    for i in range(n_dst):
        indptr[i + 1] += indptr[i]
    return indptr.astype(np.int32), src[order].astype(np.int32), order.astype(np.int32)


def first_appearance_ids(raw_ids):
    """Contiguous ids in order of first appearance (``src/builder.py:205-217``: pandas ``unique()`` order)."""
    mapping, out = {}, np.empty(len(raw_ids), dtype=np.int64)
    for i, r in enumerate(raw_ids):
        if r not in mapping:
            mapping[r] = len(mapping)
        out[i] = mapping[r]
    return out, list(mapping.keys())


def block_from_coo(num_src, num_dst, rels):
    return {'num_src': dict(num_src), 'num_dst': dict(num_dst),
            'rels': {c: (torch.as_tensor(np.asarray(s), dtype=torch.int64), torch.as_tensor(np.asarray(d), dtype=torch.int64),
                         None if w is None else torch.as_tensor(np.asarray(w), dtype=torch.float32))
                     for c, (s, d, w) in sorted(rels.items())}}


# ----------------------------------------------------------------------------------------------- model
def node_embedding(x, weight, bias):
    """src/model.py:19-24: ``nn.Linear(in_feats, out_feats)`` with bias."""
    return x @ weight.t() + bias


def neighbour_reduce(src, dst, w, x, n_dst, reducer):
    """``update_all(copy_src | u_mul_e, mean | max)`` (src/model.py:143-162, 171-208) per DGL semantics."""
    m = x[src]
    if w is not None:
        m = m * w.reshape(-1, 1).to(m.dtype)
    out = torch.zeros(n_dst, x.shape[1], dtype=x.dtype)
    if reducer == 'mean':
        out.index_add_(0, dst, m)
        deg = torch.bincount(dst, minlength=n_dst).clamp(min=1).to(x.dtype)
        return out / deg.reshape(-1, 1)
    if reducer == 'max':
        idx = dst.reshape(-1, 1).expand_as(m)
        return out.scatter_reduce(0, idx, m, reduce='amax', include_self=False)  # isolated rows stay 0
    raise KeyError(reducer)


def conv_layer(src, dst, w, h_neigh, h_self, fc_self, fc_neigh, fc_preagg, aggregator_type, norm, cetype=None):
    """src/model.py:123-237 (dropout is identity in eval mode / p=0)."""
    if aggregator_type not in AGGREGATORS:
        raise KeyError('Aggregator type {} not recognized.'.format(aggregator_type))
    use_edge = aggregator_type.endswith('_edge') and (
        cetype is None or (cetype[0] in ('user', 'item') and cetype[2] in ('user', 'item')))  # model.py:173
    base = aggregator_type[:-5] if aggregator_type.endswith('_edge') else aggregator_type
    msg = h_neigh
    if base in ('mean_nn', 'pool_nn'):
        msg = torch.relu(h_neigh @ fc_preagg.t())                                   # model.py:151,158
    red = neighbour_reduce(src, dst, w if use_edge else None, msg, h_self.shape[0],
                           'max' if base == 'pool_nn' else 'mean')
    z = torch.relu(h_self @ fc_self.t() + red @ fc_neigh.t())                        # model.py:226-227
    if norm:                                                                         # model.py:230-235
        zn = z.norm(2, 1, keepdim=True)
        zn = torch.where(zn == 0, torch.ones_like(zn), zn)
        z = z / zn
    return z


def hetero_conv(block, h, sd, prefix, aggregator_type, norm, aggregate):
    """dgl 0.5.2 ``HeteroGraphConv.forward`` on a block (restated; see oracle/dgl_shim/dgl/nn/pytorch)."""
    outputs = defaultdict(list)
    dst_inputs = {k: v[:block['num_dst'].get(k, 0)] for k, v in h.items()}
    for c in sorted(block['rels']):
        src, dst, w = block['rels'][c]
        if src.numel() == 0 or c[0] not in h or c[2] not in dst_inputs:
            continue
        p = '%s.mods.%s.' % (prefix, c[1])
        outputs[c[2]].append(conv_layer(src, dst, w, h[c[0]], dst_inputs[c[2]], sd[p + 'fc_self.weight'],
                                        sd[p + 'fc_neigh.weight'], sd.get(p + 'fc_preagg.weight'),
                                        aggregator_type, norm, c))
    out = {}
    for t, lst in outputs.items():
        st = torch.stack(lst, 0)
        if aggregate == 'sum':
            out[t] = st.sum(0)
        elif aggregate == 'mean':
            out[t] = st.mean(0)
        elif aggregate == 'max':
            out[t] = st.max(0)[0]
        else:
            raise KeyError(aggregate)
    return out


def get_repr(blocks, h, sd, aggregator_type='mean', norm=True, aggregate='sum'):
    """src/model.py:415-421."""
    for i, b in enumerate(blocks):
        h = hetero_conv(b, h, sd, 'layers.%d' % i, aggregator_type, norm, aggregate)
    return h


def embed_inputs(feats, sd):
    """src/model.py:462-466 / src/train/run.py:341-345."""
    h = {}
    for t, x in feats.items():
        key = '%s_embed.proj_feats.' % t
        h[t] = node_embedding(x, sd[key + 'weight'], sd[key + 'bias']) if key + 'weight' in sd else x
    return h


def cosine_prediction(edge_graph, h):
    """src/model.py:317-327: F.normalize(p=2, eps=1e-12) then u_dot_v; ``edge_graph``: cetype -> (src, dst)."""
    out = {}
    for c, (s, d) in sorted(edge_graph.items()):
        if c[0] not in h or c[2] not in h:
            continue
        a = torch.nn.functional.normalize(h[c[0]], p=2, dim=-1)
        b = torch.nn.functional.normalize(h[c[2]], p=2, dim=-1)
        s, d = torch.as_tensor(np.asarray(s), dtype=torch.int64), torch.as_tensor(np.asarray(d), dtype=torch.int64)
        out[c] = (a[s] * b[d]).sum(-1, keepdim=True)
    return out


def model_forward(blocks, feats, pos_edges, neg_edges, sd, aggregator_type='mean', norm=True, aggregate='sum',
                  embedding_layer=True):
    """src/model.py:423-470."""
    h = embed_inputs(feats, sd) if embedding_layer else dict(feats)
    h = get_repr(blocks, h, sd, aggregator_type, norm, aggregate)
    return h, cosine_prediction(pos_edges, h), cosine_prediction(neg_edges, h)


def max_margin_loss(pos_score, neg_score, delta, neg_sample_size, use_recency=False, recency_scores=None,
                    remove_false_negative=False, negative_mask=None):
    """src/model.py:507-533."""
    all_scores = torch.empty(0)
    for etype in pos_score.keys():
        neg = neg_score[etype].reshape(-1, neg_sample_size)
        mask = negative_mask[etype].reshape(-1, neg_sample_size) if remove_false_negative else torch.zeros(neg.shape)
        scores = torch.relu(neg + delta - pos_score[etype] - mask)
        if use_recency and recency_scores is not None and etype in recency_scores:
            scores = scores / torch.unsqueeze(recency_scores[etype], 1)
        all_scores = torch.cat((all_scores, scores.to(all_scores.dtype)), 0)
    return torch.mean(all_scores)


def get_embeddings_full(num_nodes, blocks, feats, sd, out_dim, seeds=None, aggregator_type='mean', norm=True,
                        aggregate='sum', embedding_layer=True):
    """src/train/run.py:329-349 with one full-graph block per layer: zero table per node type, rows of
    the seeded nodes overwritten with their embeddings (unseeded / unreached types stay zero)."""
    y = {t: torch.zeros(n, out_dim, dtype=next(iter(feats.values())).dtype) for t, n in num_nodes.items()}
    h = embed_inputs(feats, sd) if embedding_layer else dict(feats)
    h = get_repr(blocks, h, sd, aggregator_type, norm, aggregate)
    for t, v in h.items():
        if seeds is None:
            y[t][:] = v
        elif t in seeds:
            idx = torch.as_tensor(np.asarray(seeds[t]), dtype=torch.int64)
            y[t][idx] = v[idx]
    return y


# ----------------------------------------------------------------------------------------------- recommendation
def create_already_bought(users, items):
    """src/metrics.py:19-28 (duplicates kept, insertion order)."""
    d = defaultdict(list)
    for u, i in zip(np.asarray(users).tolist(), np.asarray(items).tolist()):
        d[u].append(i)
    return d


def cosine_scores(user_emb, item_emb, eps=1e-6):
    """``nn.CosineSimilarity(dim=1, eps=1e-6)`` (src/metrics.py:58-59), torch-1.6 formula
    ``x.y / sqrt(max(|x|^2 |y|^2, eps^2))``; identical to current torch for unit or zero rows."""
    w12 = item_emb @ user_emb
    w1 = (user_emb * user_emb).sum()
    w2 = (item_emb * item_emb).sum(1)
    return w12 / torch.sqrt(torch.clamp(w1 * w2, min=eps * eps))


def softmax(x):
    """src/utils.py:53-58."""
    e_x = np.exp(x - np.max(x))
    return e_x / e_x.sum()


def get_recs(h_user, h_item, k, user_ids, already_bought, remove_already_bought=True, popularity=None,
             weight_popularity=1):
    """src/metrics.py:52-77, one user at a time like the reference: repeat the user row, cosine against every
    item, ``np.argsort(-ratings)``, Python filter of already-bought ids, first k."""
    n_items, dim = h_item.shape
    recs = {}
    for user in user_ids:
        user_emb = h_user[user]
        bought = already_bought[user] if (hasattr(already_bought, '__missing__') or user in already_bought) else []
        rpt = torch.cat(n_items * [user_emb]).reshape(-1, dim)                       # metrics.py:55
        ratings = torch.nn.CosineSimilarity(dim=1, eps=1e-6)(rpt, h_item)            # metrics.py:58-59
        ratings = ratings.cpu().detach().numpy().reshape(n_items, )
        if popularity is not None:                                                    # metrics.py:69-72
            ratings = np.add(softmax(ratings), np.asarray(popularity).reshape(n_items, ) * weight_popularity)
        order = np.argsort(-ratings)                                                  # metrics.py:73
        if remove_already_bought:
            order = [item for item in order if item not in bought]                   # metrics.py:75
        recs[user] = order[:k]
    return recs


def recs_to_metrics(recs, ground_truth, n_items):
    """src/metrics.py:81-107: precision, recall (ground-truth duplicates counted), coverage."""
    k_rel = sum(len([i for i in iids if i in ground_truth[u]]) for u, iids in recs.items())
    k_tot = sum(len(iids) for iids in recs.values())
    r_rel = sum(len([i for i in ground_truth[u] if i in iids]) for u, iids in recs.items())
    r_tot = sum(len(ground_truth[u]) for u in recs)
    cov = len(set(i for iids in recs.values() for i in iids)) / n_items
    return k_rel / k_tot, r_rel / r_tot, cov


def get_recs_scores(h_user, h_item, user_ids):
    """fp32 rating matrix rows used by the tie-aware top-k comparison in tests (metrics.py:58-59)."""
    hu = torch.nn.functional.normalize(h_user[torch.as_tensor(np.asarray(user_ids), dtype=torch.int64)], dim=1, eps=1e-6)
    hi = torch.nn.functional.normalize(h_item, dim=1, eps=1e-6)
    return hu @ hi.t()


def get_recs_vectorised(h_user, h_item, k, user_ids, bought_indptr=None, bought_ids=None, block=4096):
    """The *fair* CPU baseline: blocked matmul + topk with the same outputs as ``get_recs`` up to ties
    (not the reference's algorithm -- reported beside it so the speed-up is not just a removed Python loop)."""
    hi = torch.nn.functional.normalize(h_item, dim=1, eps=1e-6)
    uid = np.asarray(user_ids, dtype=np.int64)
    out = np.full((uid.size, k), -1, dtype=np.int64)
    for b0 in range(0, uid.size, block):
        ub = uid[b0:b0 + block]
        r = torch.nn.functional.normalize(h_user[torch.from_numpy(ub)], dim=1, eps=1e-6) @ hi.t()
        if bought_indptr is not None:
            lo, hi_ = bought_indptr[ub], bought_indptr[ub + 1]
            rows = np.repeat(np.arange(ub.size), hi_ - lo)
            cols = np.concatenate([bought_ids[a:b] for a, b in zip(lo, hi_)]) if rows.size else np.zeros(0, np.int64)
            r[torch.from_numpy(rows), torch.from_numpy(cols.astype(np.int64))] = -float('inf')
        kk = min(k, r.shape[1])
        val, idx = torch.topk(r, kk, dim=1)
        idx = idx.numpy()
        idx[~np.isfinite(val.numpy())] = -1
        out[b0:b0 + ub.size, :kk] = idx
    return out


def merge_partial_topk(scores_list, ids_list, k):
    """Merge per-shard (score, global id) lists row-wise: keep the k best by score (ids < 0 = empty slot)."""
    s = np.concatenate(scores_list, axis=1)
    i = np.concatenate(ids_list, axis=1)
    s = np.where(i < 0, -np.inf, s)
    order = np.argsort(-s, axis=1, kind='stable')[:, :k]
    return np.take_along_axis(s, order, 1), np.take_along_axis(i, order, 1)


# ----------------------------------------------------------------------------------------------- sampled blocks
# what dgl.dataloading does behind the reference's loaders (src/sampling.py:153-207):
# MultiLayerNeighborSampler(fanouts, replace=False) / MultiLayerFullNeighborSampler frontiers, to_block compaction,
# negative_sampler.Uniform(k). DGL's own random streams cannot be reproduced (library absent, see PARITY STATUS):
# the SEMANTICS are restated (at most `fanout` distinct in-edges per seed and relation, drawn uniformly; seeds first
# in the block's source space; k negatives per positive edge laid out consecutively, destination uniform over the
# node type) on top of the product's counter-based randomness, restated here with plain Python integers.
_M64 = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15


def _fin64(x):
    x ^= x >> 30
    x = (x * 0xbf58476d1ce4e5b9) & _M64
    x ^= x >> 27
    x = (x * 0x94d049bb133111eb) & _M64
    x ^= x >> 31
    return x


def hash64(key, ctr):
    return _fin64((key + (ctr + 1) * _GOLDEN) & _M64)


def sample_key(seed, stream_id):
    return hash64(_fin64((seed + _GOLDEN) & _M64), stream_id)


def sample_frontier(indptr, indices, eperm, seeds, fanout, key, exclude=()):
    """Frontier of one relation for ``seeds`` (distinct destination ids): per seed the in-edges that are not in
    ``exclude``; with ``fanout`` >= 1 only the ``fanout`` of them with the smallest (hash64(key, eid) >> 32, CSR slot),
    emitted in CSR (= edge id) order. Returns ``(out_indptr, src_global, eids)``."""
    exclude = set(int(e) for e in exclude)
    out_indptr, src, eids = [0], [], []
    for row in seeds:
        row = int(row)
        cand = []
        for slot in range(int(indptr[row]), int(indptr[row + 1])):
            eid = int(eperm[slot]) if eperm is not None else slot
            if eid not in exclude:
                cand.append((hash64(key, eid) >> 32, slot - int(indptr[row]), eid, int(indices[slot])))
        if fanout is not None and fanout > 0 and len(cand) > fanout:
            cand = sorted(cand)[:fanout]
        cand.sort(key=lambda c: c[1])
        src += [c[3] for c in cand]
        eids += [c[2] for c in cand]
        out_indptr.append(len(src))
    return np.asarray(out_indptr, np.int32), np.asarray(src, np.int64), np.asarray(eids, np.int32)


def negative_uniform(edge_src, eids, k, n_dst_nodes, key):
    """``negative_sampler.Uniform(k)``: (src of the positive edge, uniform destination), k consecutive per edge."""
    src, dst = [], []
    for eid in eids:
        eid = int(eid)
        for j in range(k):
            src.append(int(edge_src[eid]))
            dst.append(hash64(key, eid * k + j) % n_dst_nodes)
    return np.asarray(src, np.int64), np.asarray(dst, np.int64)


def compact_block(ntypes, canonical_etypes, seeds, frontiers):
    """``to_block``: destination nodes = seeds in the given order; source nodes = the seeds first, then unseen frontier
    sources in first-appearance order (canonical etype order, then CSR order). ``frontiers[c]`` is a
    ``sample_frontier`` result. Returns ``(src_ids {nt: global ids}, rels {c: (indptr, local src, eids)})``."""
    src_ids, local = {}, {}
    for t in ntypes:
        parts = [np.asarray(seeds[t], np.int64)] if t in seeds else []
        parts += [frontiers[c][1] for c in canonical_etypes if c[0] == t and c in frontiers]
        cat = np.concatenate(parts) if parts else np.zeros(0, np.int64)
        _, uniq = first_appearance_ids([int(v) for v in cat])
        src_ids[t] = np.asarray(uniq, np.int64)
        local[t] = {g: i for i, g in enumerate(uniq)}
    rels = {}
    for c in canonical_etypes:
        if c in frontiers:
            ip, s, e = frontiers[c]
            rels[c] = (ip, np.asarray([local[c[0]][int(g)] for g in s], np.int32), e)
    return src_ids, rels
