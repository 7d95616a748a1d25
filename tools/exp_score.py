import os, sys, time, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import gnn_recsys_b200 as grb
from gnn_recsys_b200 import ops
dev = torch.device('cuda:0')
U, I, D = int(sys.argv[1]), int(sys.argv[2]), 128
parts = int(sys.argv[3]); S = int(sys.argv[4]) if len(sys.argv) > 4 else 16
g = torch.Generator(device=dev).manual_seed(0)
hu = torch.nn.functional.normalize(torch.rand(U, D, device=dev, generator=g), dim=1)
hi = torch.nn.functional.normalize(torch.rand(I, D, device=dev, generator=g), dim=1)
cfg = grb.RecsConfig(parts=parts, shortlist=S)
table = grb.ScoringTable(hi, cfg)
uq, _ = ops.score_prep(hu, None, 128, parts, cfg.elem_type, False)
for _ in range(2):
    ops.score_topk_tc(uq, table.items_q, 0, 128, parts, cfg.elem_type, None, None, S)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    ops.score_topk_tc(uq, table.items_q, 0, 128, parts, cfg.elem_type, None, None, S)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 2.0 * U * I * D * (3 if parts == 2 else 1)
print('mode=%s parts=%d S=%d U=%d I=%d: %.2f ms  executed %.0f TFLOP/s' % (os.environ.get('GR_SCORE_DEBUG_MODE', '0'), parts, S, U, I, ms, fl / ms / 1e9))
