"""Scoring-scheme sweep on a REAL workload (the bench's synthetic graph + model): per scheme the GEMM kernel time, the
whole recommend_topk time, and how many users each pass could not prove.

    python tools/exp_score.py [c1|c2|c5] [n_users_cap]
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnn_recsys_b200 as grb
from gnn_recsys_b200 import ops
import bench

name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
n_users, n_items, n_edges, n_layers, agg, hidden, out = bench.WORKLOADS[name]
dev = torch.device('cuda:0')
data = grb.make_graph_device(n_users, n_items, n_edges, seed=0, device=dev)
g = data.graph()
blk = g.full_block_on(dev)
torch.manual_seed(1)
model = grb.ConvModel(g, n_layers, {'user': 2, 'item': 4, 'hidden': hidden, 'out': out}, True, 0.0, agg).to(dev).eval()
h = model.get_repr([blk] * (n_layers - 1), model.embed({t: g.nodes[t].data['features'].to(dev) for t in g.ntypes}))
cap = int(sys.argv[2]) if len(sys.argv) > 2 else n_users
hu, hi = h['user'][:cap].contiguous(), h['item']
buys = data.relations()[('user', 'buys', 'item')]
bought = grb.BoughtCSR.from_edges(buys[0], buys[1], n_users).select(range(0, cap))
bought.on(dev)
flops = 2.0 * cap * n_items * out
rows = []
SCHEMES = [dict(), dict(shortlist=24), dict(shortlist=16), dict(k_band=False), dict(k_band=False, shortlist=16),
           dict(parts_users=2, parts_items=1, shortlist=24), dict(parts_users=2, parts_items=1, shortlist=16),
           dict(parts=2, shortlist=16), dict(parts=2, elem='bf16', shortlist=16), dict(single_cta=True)]
if len(sys.argv) > 3:
    SCHEMES = [json.loads(a) for a in sys.argv[3:]]
for kw in SCHEMES:
    cfg = grb.RecsConfig(**kw)
    table = grb.ScoringTable(hi, cfg)
    ev = {}

    def mark(nm):
        e = torch.cuda.Event(enable_timing=True); e.record(); ev[nm] = e
    for it in range(3):
        ev.clear()
        mark('t0')
        ids, sc, n_over = grb.recommend_topk(hu, table, 10, bought, return_overflow=True, mark=mark)
        mark('t1')
    torch.cuda.synchronize()
    ms = ev['score_begin'].elapsed_time(ev['score_end'])
    r = dict(cfg=kw, kernel_ms=round(ms, 3), prep_ms=round(ev['t0'].elapsed_time(ev['score_begin']), 3), total_ms=round(ev['t0'].elapsed_time(ev['t1']), 3),
             rescore_ms=round(ev['score_end'].elapsed_time(ev['rescore_end']), 3),
             second_ms=round(ev['rescore_end'].elapsed_time(ev['fallback_end']), 3), overflow=n_over,
             useful_tflops=round(flops / ms / 1e9, 1), executed_tflops=round(flops * cfg.products / ms / 1e9, 1))
    rows.append(r)
    print(json.dumps(r), flush=True)
