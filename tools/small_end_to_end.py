import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import gnn_recsys_b200 as grb
dev = torch.device('cuda:0')
data = grb.make_graph(700, 300, 9000, seed=0)
data.items[:2600] = 7  # a hub row > 2048 in-edges
g = data.graph()
torch.manual_seed(1)
for agg, nl, hid in (('mean', 2, 128), ('pool_nn', 3, 256)):
    model = grb.ConvModel(g, nl, {'user': 2, 'item': 4, 'hidden': hid, 'out': 128}, True, 0.0, agg).to(dev).eval()
    loader = grb.NodeDataLoader(g, {'user': np.arange(700), 'item': np.arange(300)}, grb.MultiLayerFullNeighborSampler(nl - 1), batch_size=None)
    y = grb.get_embeddings(g, 128, model, loader, 1, True, dev, True)
    buys = data.relations()[('user', 'buys', 'item')]
    bought = grb.BoughtCSR.from_edges(buys[0], buys[1], 700)
    for single in (False, True):
        ids = grb.get_recs_tensor(g, y, 10, np.arange(700), bought, True, dev, config=grb.RecsConfig(single_cta=single))
    ex = grb.recommend_topk(y['user'], grb.ScoringTable(y['item'], grb.RecsConfig(exact_only=True)), 10, bought)
    c = grb.metrics_from_tensor(ids, bought, 300)
raw = torch.randint(0, 500, (5000,), device=dev)
grb.ops.remap_first_appearance(raw)
u = torch.randint(0, 700, (3000,), dtype=torch.int32, device=dev); v = torch.randint(0, 300, (3000,), dtype=torch.int32, device=dev)
grb.ops.edge_cosine(u, v, y['user'], y['item'])
torch.cuda.synchronize()
print('sanitize script done')
