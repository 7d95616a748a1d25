"""One launch of the scoring kernel on synthetic clustered embeddings, sized for an ncu capture (a few waves).

    python tools/ncu_score.py [n_users] [n_items] ['{"parts": 2}'] [random|decreasing|increasing]

RecsConfig keywords as JSON. Score streams: `random` = clustered random embeddings; `decreasing` = every user's scores
fall with the item id (only the first `shortlist` items are ever inserted: the pure MMA + TMEM-drain + scan pipeline);
`increasing` = every item beats the threshold (the worst case of the insert path).
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnn_recsys_b200 as grb
from gnn_recsys_b200 import ops, recs

U = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 2
I = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
kw = json.loads(sys.argv[3]) if len(sys.argv) > 3 else {}
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
base = torch.rand(128, device=dev, generator=g)


def table(n):
    x = torch.relu(base[None, :] + 0.3 * torch.randn(n, 128, device=dev, generator=g))
    return torch.nn.functional.normalize(x, dim=1)


stream = sys.argv[4] if len(sys.argv) > 4 else 'random'
if stream == 'random':
    hu, hi = table(U), table(I)
else:
    hu = torch.zeros(U, 128, device=dev); hu[:, 0] = 1.0; hu[:, 2:] = 0.01 * torch.rand(U, 126, device=dev, generator=g)
    th = torch.linspace(0.1, 1.3, I, device=dev)
    if stream == 'increasing':
        th = th.flip(0)
    hi = torch.zeros(I, 128, device=dev); hi[:, 0] = torch.cos(th); hi[:, 1] = torch.sin(th)
cfg = grb.RecsConfig(**kw)
t = grb.ScoringTable(hi, cfg)
et = cfg.elem_type
uq, ustats = ops.score_prep(hu, None, t.d_pad, cfg.parts_users, et, True)
band = ops.score_band(t.stats, ustats, et, cfg.parts_users, cfg.parts_items, cfg.acc_err) if cfg.k_band else None
S = max(cfg.shortlist, 10)
for _ in range(2):
    ops.score_topk_tc(uq, t.items_q, 0, t.d_pad, cfg.parts_users, cfg.parts_items, et, None, None, S, 10, band, None, cfg.flags)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.score_topk_tc(uq, t.items_q, 0, t.d_pad, cfg.parts_users, cfg.parts_items, et, None, None, S, 10, band, None, cfg.flags)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps(dict(cfg=kw, stream=stream, users=U, items=I, ms=ms, useful_tflops=2.0 * U * I * 128 / ms / 1e9,
                      cycles_per_tile_at_1p9GHz=ms * 1e-3 * 1.9e9 / (((U + 255) // 256 + 147) // 148 * ((I + 127) // 128)))))
