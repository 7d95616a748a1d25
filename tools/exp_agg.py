import sys, time, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import gnn_recsys_b200 as grb
from gnn_recsys_b200 import ops
dev = torch.device('cuda:0')
U, I, E = 1_000_000, 200_000, 50_000_000
data = grb.make_graph_device(U, I, E, 0, dev)
g = data.graph()
blk = g.full_block_on(dev)
D = 128
h = {'user': torch.rand(U, D, device=dev), 'item': torch.rand(I, D, device=dev)}
w1 = torch.randn(D, D, device=dev) * 0.1; w2 = torch.randn(D, D, device=dev) * 0.1
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for c, r in blk.rels.items():
    hs, hd = h[c[0]], h[c[2]]
    out = torch.empty(r.n_dst, D, device=dev)
    agg = torch.empty(r.n_dst, D, device=dev)
    tg = timeit(lambda: ops.gather_reduce(r.indptr, r.indices, None, hs, 0, out=agg))
    tf = timeit(lambda: ops.sage_relation(r.indptr, r.indices, None, hs, hd, w1, w2, out, 0, True))
    gb = (r.nnz * (D * 4 + 4) + r.n_dst * 4) / 1e9
    print('%-12s nnz %9d n_dst %8d: gather-only %.2f ms (%.0f GB/s gather)  fused %.2f ms' % (c[1], r.nnz, r.n_dst, tg, (gb + r.n_dst*D*4/1e9) / tg * 1e3, tf))
