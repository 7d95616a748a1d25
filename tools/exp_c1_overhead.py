import sys, time, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import gnn_recsys_b200 as grb
dev = torch.device('cuda:0')
data = grb.make_graph_device(10000, 5000, 200000, 0, dev)
g = data.graph(); blk = g.full_block_on(dev)
torch.manual_seed(1)
model = grb.ConvModel(g, 2, {'user': 2, 'item': 4, 'hidden': 128, 'out': 128}).to(dev).eval()
feats = {t: g.nodes[t].data['features'].to(dev) for t in g.ntypes}
lib = grb._native.load()
for mode in (1,):
    with torch.no_grad():
        for _ in range(3):
            h = model.get_repr([blk], model.embed(dict(feats)))
        torch.cuda.synchronize()
        t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            h = model.get_repr([blk], model.embed(dict(feats)))
        e1.record(); torch.cuda.synchronize()
        print('epilogue mode %d: wall %.3f ms/iter, gpu %.3f ms/iter, launches/iter %d' % (mode, (time.perf_counter() - t0) * 50, e0.elapsed_time(e1) / 20, 0))
import cProfile, pstats
with torch.no_grad():
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20):
        h = model.get_repr([blk], model.embed(dict(feats)))
    torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
