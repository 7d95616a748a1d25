#!/usr/bin/env python
"""Benchmark of the hot path: full-graph embeddings + top-10 recommendations for every user.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c5|c4]

One "step" = one pass of the hot path over the whole synthetic graph: NodeEmbedding -> ConvLayer stack over the
four relations -> all-users x all-items cosine + top-10. Rank 0 prints ONE JSON line (see the keys at the bottom).

  value      users/s with every input already resident in HBM (features, CSR), CUDA events, max over ranks
  e2e        users/s through the public API (get_embeddings + get_recs_tensor) with HOST features (pinned), the
             H2D copy of the features and the D2H copy of the [U, 10] id table inside the timed region
  roofline   the dominant kernel (tcgen05 scoring GEMM) against the measured bf16 peak; roofline_aggregation: the
             fused CSR gather-reduce kernels against the measured HBM bandwidth
  cpu_baseline  the CPU oracle (reference semantics) on the box's host cores, bounded sample (rank 0, N = 1 only)

--impl reference times the reference's own algorithm (oracle/straightline.py: the per-user get_recs loop of
src/metrics.py:52-77 and the layer-wise CPU embedding pass) on the host cores. DGL 0.5.2 is not installable in this
image, so the reference arm is the oracle PORT (kind "port"); nothing of the CUDA engine runs on that arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_RECS = 10
WORKLOADS = {  # BASELINE.json configs: users, items, edges, n_layers, aggregator, hidden, out
    'c1': (10_000, 5_000, 200_000, 2, 'mean', 128, 128),
    'c2': (1_000_000, 200_000, 50_000_000, 2, 'mean', 128, 128),
    'c3': (5_000_000, 500_000, 200_000_000, 3, 'pool_nn', 256, 128),
    'c5': (10_000_000, 1_000_000, 500_000_000, 2, 'mean', 128, 128),
    # training-step forward (BASELINE configs[3]) on the c1 graph: fan-out [10, 10] blocks, 1024 positive edges + K
    # uniform negatives each, CosinePrediction + max-margin loss. A side measurement (its own metric), not the bench line.
    'c4': (10_000, 5_000, 200_000, 3, 'mean', 128, 128),
}
C4_BATCH, C4_FANOUTS, C4_DELTA = 1024, [10, 10], 0.266
REV_ETYPES = {'buys': 'bought-by', 'bought-by': 'buys', 'clicks': 'clicked-by', 'clicked-by': 'clicks'}


def dram_traffic(config, stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per step of one stage's kernels, from the committed ncu capture of
    THIS round's kernels: profiles/traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` raw CSV.
    None when no capture of the current kernels exists for the config (never a stale constant)."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    e = t.get(config, {}).get(stage)
    return None if e is None or e.get('partial') else e.get('dram_bytes')


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        cl = d.get('clocks_under_load') or {}
        return dict(hbm=d['hbm_gbs'], tc_burst=d['bf16_tflops'], tc=d['bf16_tflops_sustained'], source='measured',
                    sustained_mhz=cl.get('sm_mhz_median'), max_mhz=d.get('sm_max_mhz'))
    return dict(hbm=6650.0, tc_burst=1590.0, tc=1400.0, source='fallback', sustained_mhz=None, max_mhz=None)


def tensor_peak(pk, clocks):
    """(peak TFLOP/s, label): the BURST figure when the timed region ran at burst clocks (a step too short to reach the
    power cap: median SM clock nearer the maximum than the clock the sustained figure was measured at), else the
    SUSTAINED one -- MEASURED_PEAKS.json holds both regimes of the same cuBLAS bf16 GEMM."""
    mhz = (clocks or {}).get('sm_mhz')
    hi = pk.get('max_mhz') or (clocks or {}).get('sm_max_mhz')
    lo = pk.get('sustained_mhz') or (0.7 * hi if hi else None)
    if mhz and hi and lo and mhz >= 0.5 * (lo + hi):
        return pk['tc_burst'], pk['source'] + ' (burst: %d MHz median under load, sustained figure measured at %d MHz)' % (mhz, lo)
    return pk['tc'], pk['source'] + ' (sustained)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln.split(', ') for ts, ln in self.lines if t0 <= ts <= t1 + 0.15] or [ln.split(', ') for _, ln in self.lines]
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.strip().lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_reference_sample(n_users, n_items, n_edges, d, budget_s=25.0, max_users=200, seed=0, vectorised=True):
    """CPU baselines on a bounded sample, scaled to users/s of the whole job:
    (a) layer-wise embedding pass (torch CPU index_add_ / scatter: already vectorised, like DGL's C++ SpMM) on a c1-sized
        graph, scaled by edge count;
    (b) REFERENCE semantics: the per-user get_recs loop of src/metrics.py:52-77 (oracle port) for up to `max_users` users
        against a full-size item table -> `value`;
    (c) FAIR vectorised path: blocked matmul + topk + bought filter (oracle.get_recs_vectorised) on all cores for a few
        thousand users -> `vectorised`, so that the speed-up is not merely "Python loop removed"."""
    from oracle import straightline as O
    import gnn_recsys_b200 as grb
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    # (a) embeddings on a small graph of the same shape family
    su, si, se = 10_000, 5_000, 200_000
    data = grb.make_graph(su, si, se, seed)
    num = {'user': su, 'item': si}
    blk = O.block_from_coo(num, num, {c: (s.astype(np.int64), t.astype(np.int64), None) for c, (s, t) in data.relations().items()})
    g = torch.Generator().manual_seed(seed + 1)
    sd = {}
    for t, f in (('user', 2), ('item', 4)):
        sd['%s_embed.proj_feats.weight' % t] = torch.randn(d, f, generator=g) * 0.5
        sd['%s_embed.proj_feats.bias' % t] = torch.randn(d, generator=g) * 0.1
    for et in ('buys', 'bought-by', 'clicks', 'clicked-by'):
        for nm in ('fc_self', 'fc_neigh'):
            sd['layers.0.mods.%s.%s.weight' % (et, nm)] = torch.randn(d, d, generator=g) * (2.0 / d) ** 0.5
    feats = {'user': data.user_feat, 'item': data.item_feat}
    t0 = time.perf_counter()
    y = O.get_embeddings_full(num, [blk], feats, sd, d)
    t_embed_small = time.perf_counter() - t0
    t_embed_est = t_embed_small * (n_edges / se)
    # (b) per-user recommendation loop against a full-size item table (cost is data-independent)
    reps = (n_items + si - 1) // si
    h_item = y['item'].repeat(reps, 1)[:n_items].contiguous()
    h_user = y['user']
    buys = data.relations()[('user', 'buys', 'item')]
    bought = O.create_already_bought(buys[0], buys[1])
    done, t_recs = 0, 0.0
    while done < max_users and t_recs < budget_s:
        t0 = time.perf_counter()
        O.get_recs(h_user, h_item, K_RECS, [done], bought)
        t_recs += time.perf_counter() - t0
        done += 1
    per_user = t_recs / done
    users_per_s = n_users / (t_embed_est + n_users * per_user)
    sample = ('oracle port of src/metrics.py:52-77 get_recs loop: %d users x %d items (%.3f s/user) + layer-wise CPU embedding '
              'pass on a %dx%dx%d graph (%.2f s) scaled by edge count to %.0f s; users/s = U / (t_embed + U * t_user)'
              % (done, n_items, per_user, su, si, se, t_embed_small, t_embed_est))
    out = dict(value=users_per_s, unit='users/s', cores=cores, kind='port', sample=sample, s_per_user=per_user,
               embed_s_est=t_embed_est)
    if vectorised:  # (c)
        bc = grb.BoughtCSR.from_edges(buys[0], buys[1], su)
        nv = int(min(su, max(256, 2_000_000_000 // max(n_items, 1))))  # ~2e9 scores per timed call
        uids = np.arange(nv)
        O.get_recs_vectorised(h_user, h_item, K_RECS, uids[:64], bc.indptr, bc.ids.astype(np.int64))   # warm-up
        t0 = time.perf_counter()
        O.get_recs_vectorised(h_user, h_item, K_RECS, uids, bc.indptr, bc.ids.astype(np.int64))
        t_vec = (time.perf_counter() - t0) / nv
        out['vectorised'] = dict(value=n_users / (t_embed_est + n_users * t_vec), unit='users/s', cores=cores, s_per_user=t_vec,
                                 sample='blocked torch matmul + topk + bought filter (oracle.get_recs_vectorised, all cores): '
                                        '%d users x %d items (%.2e s/user) + the same embedding estimate' % (nv, n_items, t_vec))
    return out


def run_reference_arm(args, wl_name, wl):
    n_users, n_items, n_edges, n_layers, agg, hidden, out = wl
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vals = []
    base = None
    t_start = time.perf_counter()
    for step in range(args.warmup + args.steps):
        base = cpu_reference_sample(n_users, n_items, n_edges, out, budget_s=8.0, max_users=32, seed=step, vectorised=False)
        if step >= args.warmup:
            vals.append(base['value'])
    v = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': 'users/sec for full-graph embed+top-10 recs', 'value': v, 'unit': 'users/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * (time.perf_counter() - t_start) / max(1, args.steps + args.warmup),
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(wl_name, wl, args.gpus),
        'cpu_baseline': dict(base, value=v),
        'e2e': {'value': v, 'unit': 'users/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(name, wl, n_gpus, item_shards=None):
    n_users, n_items, n_edges, n_layers, agg, hidden, out = wl
    return {'workload': '%s: %d users x %d items x %d click/purchase edges, %d-layer ConvModel %s, hidden %d / out %d, '
                        'full-graph embeddings + top-%d recs for every user' % (name, n_users, n_items, n_edges, n_layers,
                                                                                 agg, hidden, out, K_RECS),
            'users': n_users, 'items': n_items, 'edges': n_edges, 'k': K_RECS,
            'parallelism': 'single GPU' if n_gpus == 1 else (
                ('contiguous id-range shards x%d: CSR rows, features and NodeEmbedding per rank, NCCL all-gather per layer; '
                 'scoring: ' % n_gpus) +
                ('item-range shards, per-shard top-k merged on the rank owning the user range (all-to-all)' if item_shards != 1
                 else 'user-range shards against the all-gathered item table')),
            'l2': 'inputs larger than L2 (CSR + tables >> 126 MB per step), no explicit flush'}



# ------------------------------------------------------------------------------------------------ config 4
def c4_oracle_step(g, batch, sd, neg_k):
    """One training-step forward on the CPU oracle (reference semantics, torch CPU) for a host-built batch."""
    from oracle import straightline as O
    _, pos_g, neg_g, blocks = batch
    obs = []
    for b in blocks:
        rels = {}
        for c, r in b.rels.items():
            dst = np.repeat(np.arange(r.n_dst), np.diff(r.indptr.numpy()))
            rels[c] = (r.indices.numpy().astype(np.int64), dst, None)
        obs.append(O.block_from_coo(b.num_src, b.num_dst, rels))
    feats = {t: v for t, v in blocks[0].srcdata['features'].items()}
    pe = {c: pos_g.edge_arrays(c) for c in pos_g.canonical_etypes}
    ne = {c: neg_g.edge_arrays(c) for c in neg_g.canonical_etypes}
    t0 = time.perf_counter()
    h, pos, neg = O.model_forward(obs, feats, pe, ne, sd)
    loss = O.max_margin_loss(pos, neg, C4_DELTA, neg_k)
    return time.perf_counter() - t0, float(loss)


def run_c4(args, wl):
    """BASELINE configs[3]: EdgeDataLoader blocks (fan-out [10, 10]) + ConvModel.forward + max_margin_loss.
    `value`: blocks built ON THE DEVICE (device= loader), CUDA events. `e2e`: the reference's call shape -- host loader,
    block.to(device), forward, loss.item(). cpu_baseline / --impl reference: the oracle on the host cores."""
    import gnn_recsys_b200 as grb
    n_users, n_items, n_edges, n_layers, agg, hidden, out = wl
    k = args.neg_k
    data = grb.make_graph(n_users, n_items, n_edges, 0)
    g = data.graph()
    eids = {'buys': np.arange(g.num_edges('buys')), 'clicks': np.arange(g.num_edges('clicks'))}
    kw = dict(exclude='reverse_types', reverse_etypes=REV_ETYPES, negative_sampler=grb.negative_sampler.Uniform(k),
              batch_size=C4_BATCH, shuffle=True, seed=2)
    workload = ('c4: training-step forward on %d users x %d items x %d edges: EdgeDataLoader fan-out %s, %d positive + %d x %d '
                'negative edges, %d-layer ConvModel %s %d/%d, CosinePrediction + max-margin loss'
                % (n_users, n_items, n_edges, C4_FANOUTS, C4_BATCH, C4_BATCH, k, n_layers, agg, hidden, out))
    metric = 'training-step forwards/sec (sampled blocks + positive/negative cosine scores + loss)'
    torch.manual_seed(1)
    if args.impl == 'reference':
        if int(os.environ.get('RANK', '0')) != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        stub = grb.ConvModel(g, n_layers, {'user': 2, 'item': 4, 'hidden': hidden, 'out': out}, True, 0.0, agg)
        sd = {kk: v.detach() for kk, v in stub.state_dict().items()}
        it = iter(grb.EdgeDataLoader(g, eids, grb.MultiLayerNeighborSampler(C4_FANOUTS), **kw))
        ts = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            batch = next(it)
            c4_oracle_step(g, batch, sd, k)
            if i >= args.warmup:
                ts.append(time.perf_counter() - t0)
        v = 1.0 / float(np.mean(ts))
        cores = torch.get_num_threads()
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': v, 'unit': 'steps/s', 'n_gpus': args.gpus,
                          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 / v, 'higher_is_better': True,
                          'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': workload},
                          'cpu_baseline': {'value': v, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
                                           'sample': 'host block builder + oracle model_forward + loss, %d steps' % args.steps},
                          'e2e': {'value': v, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                          'gpu_launches': 0}))
        return
    N = grb._native
    N.load()
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    model = grb.ConvModel(g, n_layers, {'user': 2, 'item': 4, 'hidden': hidden, 'out': out}, True, 0.0, agg).to(dev).eval()
    pk = peaks()
    g.full_block_on(dev)
    dev_loader = grb.EdgeDataLoader(g, eids, grb.MultiLayerNeighborSampler(C4_FANOUTS), device=dev, **kw)
    host_loader = grb.EdgeDataLoader(g, eids, grb.MultiLayerNeighborSampler(C4_FANOUTS), **kw)
    neg_ev = []

    def forward(batch, record=False):
        _, pos_g, neg_g, blocks = batch
        blocks = [b.to(dev) for b in blocks]
        h = {t: v.to(dev, torch.float32, non_blocking=True) for t, v in blocks[0].srcdata['features'].items()}
        h = model.get_repr(blocks, model.embed(h))
        pos = model.pred_fn(pos_g, h)
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        neg = model.pred_fn(neg_g, h)
        if record:
            e1.record()
            neg_ev.append((e0, e1, sum(int(v.shape[0]) for v in neg.values())))
        return grb.max_margin_loss(pos, neg, C4_DELTA, k, cuda=True, device=dev)

    sampler = ClockSampler(dev.index or 0)
    it = iter(dev_loader)
    for _ in range(max(args.warmup, 3)):
        loss = forward(next(it))
    torch.cuda.synchronize()
    launches0 = N.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        loss = forward(next(it), record=True)
    ev1.record()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    launches = N.kernel_launches() - launches0
    clocks = sampler.stop(w0, w1)
    ms_step = ev0.elapsed_time(ev1) / args.steps
    # forward alone on a resident batch (what remains once block building is off the critical path)
    batch = next(it)
    for _ in range(3):
        forward(batch)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        forward(batch)
    ev1.record()
    torch.cuda.synchronize()
    ms_fwd = ev0.elapsed_time(ev1) / args.steps
    # reference call shape: host loader -> .to(device) -> forward -> loss.item()
    hit = iter(host_loader)
    for _ in range(2):
        float(forward(next(hit)))
    h2d = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b = next(hit)
        h2d += sum(r.csr_bytes() for blk in b[3] for r in blk.rels.values())
        h2d += sum(int(v.numel()) * 4 for v in b[3][0].srcdata['features'].values())
        h2d += sum(16 * gph.num_edges(c) for gph in (b[1], b[2]) for c in gph.canonical_etypes)
        float(forward(b))
    e2e_s = (time.perf_counter() - t0) / args.steps
    # edge-scoring kernel roofline: 8*D + 12 algorithmic bytes per edge
    ms_neg = float(np.mean([a.elapsed_time(b) for a, b, _ in neg_ev]))
    n_neg = int(np.mean([n for _, _, n in neg_ev]))
    gbs = n_neg * (8 * out + 12) / (ms_neg * 1e-3) / 1e9
    roof = {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': gbs / pk['hbm'], 'traffic': None,
            'kernel': 'edge_cosine_kernel over the negative edges (both etypes, 2 launches)', 'of': pk['source'],
            'note': 'gathers hit L2: the batch embedding tables are a few MB; algorithmic bytes = (8*D + 12) per edge'}
    cpu = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {kk: v.detach().cpu() for kk, v in model.state_dict().items()}
        hb = next(iter(grb.EdgeDataLoader(g, eids, grb.MultiLayerNeighborSampler(C4_FANOUTS), **kw)))
        ts = [c4_oracle_step(g, hb, sd, k)[0] for _ in range(3)]
        cpu = {'value': 1.0 / min(ts), 'unit': 'steps/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': 'oracle model_forward + loss on one host-built batch (blocks given), best of 3'}
    print(json.dumps({
        'metric': metric, 'value': 1e3 / ms_step, 'unit': 'steps/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': workload, 'l2': 'working set of a batch fits L2; no flush (latency-bound step)'},
        'e2e': {'value': 1.0 / e2e_s, 'unit': 'steps/s', 'h2d_bytes_per_step': h2d // args.steps, 'd2h_bytes_per_step': 4,
                'ms_per_step': e2e_s * 1e3, 'note': 'host (NumPy) block builder + block.to(device) + forward + loss.item()'},
        'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
        'stages_ms': {'device_blocks_plus_forward_ms': ms_step, 'forward_only_ms': ms_fwd, 'negative_scoring_ms': ms_neg},
        'scored_edges_per_step': C4_BATCH * (1 + k), 'loss': float(loss)}))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--elem', default='fp16', choices=['bf16', 'fp16'])
    ap.add_argument('--parts', type=int, default=None, choices=[1, 2], help='shorthand: 1 = single product, 2 = 3-product hi/lo split without a second pass')
    ap.add_argument('--parts-users', type=int, default=1, choices=[1, 2])
    ap.add_argument('--parts-items', type=int, default=1, choices=[1, 2])
    ap.add_argument('--shortlist', type=int, default=32)
    ap.add_argument('--no-k-band', action='store_true')
    ap.add_argument('--no-item-order', action='store_true', help='sweep the items in id order (A/B of RecsConfig.item_order)')
    ap.add_argument('--item-shards', type=int, default=None, help='N>1 scoring layout: N = item-range shards + owner-side top-k merge; 1 = user-range shards, replicated item table; default: shard the longer side')
    ap.add_argument('--neg-k', type=int, default=2500, help='c4: negatives per positive edge (reference default 2500)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-verify', action='store_true')
    ap.add_argument('--one-layout', action='store_true', help='N>1: time only the primary scoring layout')
    ap.add_argument('--equal-rows', action='store_true', help='N>1: cut the item id space by rows instead of by work')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = WORKLOADS[args.config]
    if args.config == 'c4':
        return run_c4(args, wl)
    if args.impl == 'reference':
        return run_reference_arm(args, args.config, wl)

    import torch.distributed as dist
    import gnn_recsys_b200 as grb
    from gnn_recsys_b200 import distributed as D
    N = grb._native
    N.load()  # fails loudly when the CUDA extension is missing

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_users, n_items, n_edges, n_layers, agg, hidden, out = wl
    if args.item_shards is None and world > 1:
        args.item_shards = D.choose_item_shards(n_users, n_items, world)
    pk = peaks()

    # ---- synthetic data (same on every rank: seeded), device CSR, model
    t0 = time.perf_counter()
    data = grb.make_graph_device(n_users, n_items, n_edges, seed=0, device=dev)
    g = data.graph()
    num = {'user': n_users, 'item': n_items}
    t_gen = time.perf_counter() - t0
    torch.manual_seed(1)
    model = grb.ConvModel(g, n_layers, {'user': 2, 'item': 4, 'hidden': hidden, 'out': out}, True, 0.0, agg, 'cos',
                          'sum', True).to(dev).eval()
    n_conv = n_layers - 1
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    mem0 = torch.cuda.memory_allocated()
    t0 = time.perf_counter()
    if world == 1:
        ranges = {t: (0, n) for t, n in num.items()}
        blk = g.full_block_on(dev)
    else:  # SHARDED STORAGE: this rank ingests and keeps only the CSR rows / feature rows of its destination id ranges
        # (users: equal row ranges; items: ranges of equal WORK -- the most popular item alone holds ~8 % of the edges)
        wb = {} if args.equal_rows else {'item': D.work_bounds(g, 'item', world)}
        ranges = D.node_ranges(num, world, rank, wb)
        blk = g.sharded_block_on(dev, ranges, bounds=wb)
    torch.cuda.synchronize()
    t_ingest = time.perf_counter() - t0
    blocks = [blk] * n_conv
    feats_host = {t: g.nodes[t].data['features'][ranges[t][0]:ranges[t][1]].contiguous().pin_memory() for t in g.ntypes}
    feats_dev = {t: v.to(dev) for t, v in feats_host.items()}
    buys = data.relations()[('user', 'buys', 'item')]
    bought = grb.BoughtCSR.from_edges(buys[0], buys[1], n_users)
    bought.on(dev)
    graph_bytes = sum(r.indptr.numel() * 4 + r.indices.numel() * 4 + (r.eperm.numel() * 4 if r.eperm is not None else 0)
                      for r in blk.rels.values())
    resident = {'graph_csr_bytes': int(graph_bytes), 'feature_bytes': int(sum(v.numel() * 4 for v in feats_dev.values())),
                'allocated_after_ingest_bytes': int(torch.cuda.memory_allocated() - mem0)}
    cfg = grb.RecsConfig(elem=args.elem, parts=args.parts, parts_users=args.parts_users, parts_items=args.parts_items,
                         shortlist=args.shortlist, k_band=not args.no_k_band, item_order=not args.no_item_order)
    uid_all = np.arange(n_users)
    layouts = ['single'] if world == 1 else (['user_shards', 'item_shards'] if args.item_shards == 1 else ['item_shards', 'user_shards'])
    if args.one_layout:
        layouts = layouts[:1]
    own_range = {}  # layout -> user range whose recommendations this rank returns

    def forward_pass(layout, feats, mark):
        """NodeEmbedding + conv layers. world > 1: own rows only, all-gather per layer (distributed.sharded_forward)."""
        if world == 1:
            h = model.embed(dict(feats))
            mark('embed_in')
            h = model.get_repr(blocks, h)
            mark('aggregate')
            return h
        # user_shards scores this rank's users against ALL items: only the item table is gathered after the last layer.
        # item_shards needs all users and its own (equal-row) item range; the aggregation ranges of the items are cut
        # by work, not rows, so the (small) item table is gathered too
        h = D.sharded_forward(model, blocks, feats, gather_last=('item',) if layout == 'user_shards' else None, mark=mark)
        mark('aggregate')
        return h

    def score_pass(layout, h, mark):
        if world == 1:
            table = grb.ScoringTable(h['item'], cfg)
            ids, scores, n_over = grb.recommend_topk(h['user'], table, K_RECS, bought, return_overflow=True, mark=mark)
            own_range[layout] = (0, n_users)
            return ids, n_over
        ids, scores, owned, n_over = D.sharded_recommend(h['user'], h['item'], K_RECS, bought, cfg, mark=mark,
                                                         item_shards=1 if layout == 'user_shards' else world,
                                                         return_overflow=True)
        own_range[layout] = owned
        return ids, n_over

    def resident_step(layout, record=None):
        """Inputs resident in HBM. Returns (ids, n_overflow, h); `record` collects CUDA events per stage."""
        ev = {}

        def mark(name):
            if record is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev[name] = e
        mark('t0')
        h = forward_pass(layout, feats_dev, mark)
        ids, n_over = score_pass(layout, h, mark)
        mark('t1')
        if record is not None:
            record.append(ev)
        return ids, n_over, h

    loader = grb.NodeDataLoader(g, {'user': uid_all, 'item': np.arange(n_items)},
                                grb.MultiLayerFullNeighborSampler(n_conv), batch_size=None)
    ids_pinned = torch.empty((D.chunk_rows(n_users, world), K_RECS), dtype=torch.int32).pin_memory()
    if world == 1:
        for t in g.ntypes:
            g.nodes[t].data['features'] = feats_host[t]   # pinned: the public API copies them with non_blocking=True

    def e2e_step(layout):
        """Public API with HOST features (pinned): H2D of the features and D2H of the id table inside the call."""
        if world == 1:
            y = grb.get_embeddings(g, out, model, loader, 1, True, dev, True)
            ids = grb.get_recs_tensor(g, y, K_RECS, uid_all, bought, True, dev, config=cfg)
        else:
            feats = {t: v.to(dev, non_blocking=True) for t, v in feats_host.items()}   # this rank's rows only
            h = forward_pass(layout, feats, lambda name: None)
            ids, _ = score_pass(layout, h, None)
        host_ids = ids_pinned[:ids.shape[0]]
        host_ids.copy_(ids, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_ids

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_true(flag):
        if world > 1:
            t = torch.tensor([1.0 if flag else 0.0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item() > 0.5)
        return bool(flag)

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(x)

    # nvidia-smi starts BEFORE the warm-up: its start-up (NVML init) stalls kernel launches for milliseconds, which
    # must not land in the timed region (it showed up as a 6 ms "aggregate" stage on the 2 ms c1 step)
    sampler = ClockSampler(local) if rank == 0 else None

    # ---- warm-up (also the verification pass)
    last = {}
    for layout in layouts:
        for _ in range(max(args.warmup, 1)):
            last[layout] = resident_step(layout)
    barrier()
    resident['peak_allocated_bytes'] = int(torch.cuda.max_memory_allocated())
    verified, verified_emb = {}, None
    if not args.no_verify:
        for layout in layouts:
            ids, n_over, h = last[layout]
            # 512 sampled users of the range this rank owns, against the brute-force fp32 kernel over ALL items
            ub, ue = own_range[layout]
            h_item_full = h['item']
            sample = torch.from_numpy(ub + np.random.default_rng(rank).choice(ue - ub, min(512, ue - ub), replace=False)).to(dev)
            ex_tab = grb.ScoringTable(h_item_full, grb.RecsConfig(exact_only=True))
            ex_ids, ex_sc = grb.recommend_topk(h['user'][sample], ex_tab, K_RECS, bought.select(sample.cpu().numpy()))
            hi_n = torch.nn.functional.normalize(h_item_full, dim=1)
            hu_n = torch.nn.functional.normalize(h['user'][sample], dim=1)
            mine = ids[sample - ub]
            s_got = (hu_n.unsqueeze(1) * hi_n[mine.long().clamp(min=0)]).sum(-1)
            s_ex = (hu_n.unsqueeze(1) * hi_n[ex_ids.long().clamp(min=0)]).sum(-1)
            ok = bool(((s_got - s_ex).abs() < 1e-5).all()) and bool(((mine < 0) == (ex_ids < 0)).all())
            verified[layout] = all_true(ok)
        if world > 1:
            # sharded storage == the un-sharded pass: this rank's embedding rows (both node types) against
            # model.get_repr over the FULL graph block, built here for the check only and dropped afterwards
            h_sh = D.sharded_forward(model, blocks, feats_dev)
            full_blk = g.full_block_on(dev)
            feats_full = {t: g.nodes[t].data['features'].to(dev) for t in g.ntypes}
            h_full = model.get_repr([full_blk] * n_conv, model.embed(feats_full))
            ok = True
            for t in ('user', 'item'):
                b, e = ranges[t]
                rows = torch.from_numpy(b + np.random.default_rng(rank + 7).choice(e - b, min(4096, e - b), replace=False)).to(dev)
                other = torch.from_numpy(np.random.default_rng(rank + 9).choice(num[t], min(4096, num[t]), replace=False)).to(dev)
                ok = ok and bool(torch.equal(h_sh[t][rows], h_full[t][rows])) and bool(torch.equal(h_sh[t][other], h_full[t][other]))
            verified_emb = all_true(ok)
            del h_sh, h_full, feats_full, full_blk
            g._dev_blocks = {k_: v_ for k_, v_ in g._dev_blocks.items() if v_ is blk}
            torch.cuda.empty_cache()
    del last

    # ---- timed: resident inputs, CUDA events, max over ranks (primary layout = `value`; the other one after it)
    timed = {}
    launches0 = N.kernel_launches()
    clocks = None
    for li, layout in enumerate(layouts):
        rec = []
        barrier()
        w0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            ids, n_over, h = resident_step(layout, rec)
        ev1.record()
        barrier()
        w1 = time.perf_counter()
        if li == 0:
            launches = N.kernel_launches() - launches0
            clocks = sampler.stop(w0, w1) if sampler is not None else None
        ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps

        def stage_ms(a, b_, rec=rec):
            ok = rec and a in rec[0] and b_ in rec[0]
            return float(np.mean([e[a].elapsed_time(e[b_]) for e in rec])) if ok else None
        stages = {'embed_in_ms': stage_ms('t0', 'embed_in'), 'aggregate_ms': stage_ms('embed_in', 'aggregate'),
                  'prep_ms': stage_ms('aggregate', 'score_begin'), 'score_ms': stage_ms('score_begin', 'score_end'),
                  'rescore_ms': stage_ms('score_end', 'rescore_end'), 'second_pass_ms': stage_ms('rescore_end', 'fallback_end'),
                  'rescore_merge_ms': stage_ms('score_end', 't1')}
        if world > 1:
            comp = [stage_ms('embed_in' if i == 0 else 'gather%d' % (i - 1), 'compute%d' % i) for i in range(n_conv)]
            gath = [stage_ms('compute%d' % i, 'gather%d' % i) for i in range(n_conv)]
            stages['aggregate_kernels_ms'] = float(sum(comp)) if all(c is not None for c in comp) else None
            stages['aggregate_allgather_ms'] = float(sum(gath)) if all(c is not None for c in gath) else None
            stages['note'] = ('rank 0; embed_in_ms = NodeEmbedding of own rows + all-gather of the embedded inputs; aggregate_ms = '
                              'kernels + per-layer NCCL all-gather; rescore_merge_ms includes the all-to-all + merge (item_shards)')
        timed[layout] = dict(ms_step=ms_step, value=n_users / (ms_step * 1e-3), stages=stages, n_over=n_over)
    primary = layouts[0]
    ms_step, value, stages, n_over = (timed[primary][k_] for k_ in ('ms_step', 'value', 'stages', 'n_over'))

    # ---- timed: end to end through the public API with host buffers (primary layout)
    for _ in range(min(args.warmup, 2)):
        e2e_step(primary)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids_host = e2e_step(primary)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    h2d = sum(int(v.numel()) * 4 for v in feats_host.values())   # this rank's feature rows ...
    d2h = int(ids_host.numel()) * 4                              # ... and its users' id table
    if world > 1:  # whole-job bytes: summed over the ranks
        t = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d, d2h = int(t[0].item()), int(t[1].item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines
    D_ = hidden
    e_into = {'item': int(data.is_buy.sum()) + int((~data.is_buy).sum()), 'user': n_edges}
    agg_bytes = 0
    for t, n_dst in (('item', n_items), ('user', n_users)):
        agg_bytes += 2 * 4 * (n_dst + 1) + 4 * n_edges + 4 * D_ * n_edges + 4 * D_ * n_dst + 4 * out * n_dst
    agg_bytes *= n_conv
    roof_agg = None
    agg_ms = stages.get('aggregate_kernels_ms') if world > 1 else stages.get('aggregate_ms')
    if agg_ms:
        gbs = agg_bytes / world / (agg_ms * 1e-3) / 1e9
        traffic = dram_traffic(args.config, 'aggregate') if world == 1 else None
        roof_agg = {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': gbs / pk['hbm'],
                    'traffic': traffic, 'dram_gbs': (traffic / (agg_ms * 1e-3) / 1e9) if traffic else None,
                    'algorithmic_bytes': agg_bytes // world, 'of': pk['source'], 'per': 'GPU', 'kernel_ms': agg_ms,
                    'note': 'all fused CSR gather-reduce + projection kernels of one step (4 relations%s): kernel time only; '
                            'achieved = ALGORITHMIC bytes (SURVEY 8d B_dst, neighbour rows counted per edge) / time; '
                            'dram_gbs = measured DRAM bytes of the committed ncu capture / time (popular rows hit L2)'
                            % ('' if world == 1 else ", this rank's 1/%d of the rows; the all-gather is in aggregate_allgather_ms" % world)}
        if world > 1 and stages.get('aggregate_ms'):
            roof_agg['with_allgather_gbs'] = agg_bytes / world / (stages['aggregate_ms'] * 1e-3) / 1e9
    flops = 2.0 * n_users * n_items * out
    roof = None
    if stages.get('score_ms'):
        tf = flops / world / (stages['score_ms'] * 1e-3) / 1e12
        tc_peak, tc_label = tensor_peak(pk, clocks)
        roof = {'bound': 'tensor', 'achieved': tf, 'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': tf / tc_peak,
                'traffic': dram_traffic(args.config, 'score') if world == 1 else None, 'per': 'GPU',
                'executed_tflops': tf * cfg.products, 'executed_frac': tf * cfg.products / tc_peak,
                'of': tc_label,
                'kernel': 'score_topk_kernel (tcgen05 kind::f16 %s, %d-product first pass, fused top-%d shortlist)' % (args.elem, cfg.products, args.shortlist),
                'note': 'achieved counts 2*U*I*D once over the first-pass kernel time (score_ms); that pass executes %dx of '
                        'it on the tensor pipe; users it cannot prove (overflow_users[0]) go through the %s second pass, '
                        'timed in second_pass_ms' % (cfg.products, '3-product' if cfg.second else 'exact fp32')}
    elif roof_agg is not None:
        roof = roof_agg

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_reference_sample(n_users, n_items, n_edges, out)

    line = {
        'metric': 'users/sec for full-graph embed+top-10 recs', 'value': value, 'unit': 'users/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32 embeddings + %s x%d-product tcgen05 scoring (f32 accumulate, exact f32 re-score + proof%s)' % (args.elem, cfg.products, ', 3-product second pass' if cfg.second else ''),
        'data': 'synthetic', 'config': workload_config(args.config, wl, world, None if world == 1 else (1 if primary == 'user_shards' else world)),
        'e2e': {'value': n_users / e2e_s, 'unit': 'users/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_s * 1e3},
        'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'roofline_aggregation': roof_agg,
        'cpu_baseline': cpu, 'stages_ms': stages, 'overflow_users': list(n_over),
        'verified_vs_exact_fp32': (all(verified.values()) if verified else None), 'verified_layouts': verified or None,
        'verified_sharded_embeddings': verified_emb,
        'layouts': {k_: {'users_per_s': v_['value'], 'ms_per_step': v_['ms_step'], 'stages_ms': v_['stages']} for k_, v_ in timed.items()},
        'resident_per_rank': resident,
        'setup_s': {'generate': t_gen, 'ingest_csr_build': t_ingest},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
