import sys, numpy as np, torch, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import gnn_recsys_b200 as grb
from oracle import straightline as O
torch.manual_seed(1)
U,I,E = int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3]); D=128
d = grb.make_graph(U,I,E,0); rel = d.relations(); num={'user':U,'item':I}
sd={}
for t,f in (('user',2),('item',4)):
    l=torch.nn.Linear(f,D); sd['%s_embed.proj_feats.weight'%t]=l.weight.detach(); sd['%s_embed.proj_feats.bias'%t]=l.bias.detach()
gain=torch.nn.init.calculate_gain('relu')
for et in ('buys','bought-by','clicks','clicked-by'):
    for nm in ('fc_self','fc_neigh'):
        w=torch.empty(D,D); torch.nn.init.xavier_uniform_(w,gain=gain); sd['layers.0.mods.%s.%s.weight'%(et,nm)]=w
blk = O.block_from_coo(num,num,{c:(s.astype(np.int64),t.astype(np.int64),None) for c,(s,t) in rel.items()})
y=O.get_embeddings_full(num,[blk],{'user':d.user_feat,'item':d.item_feat},sd,D)
hu=torch.nn.functional.normalize(y['user'],dim=1); hi=torch.nn.functional.normalize(y['item'],dim=1)
def kmeans_order(x, k, iters=8, seed=0):
    g=torch.Generator().manual_seed(seed)
    c = x[torch.randperm(x.shape[0],generator=g)[:k]].clone()
    for _ in range(iters):
        a = (x@c.t()).argmax(1)
        for j in range(k):
            m = a==j
            if m.any(): c[j]=torch.nn.functional.normalize(x[m].mean(0),dim=0)
    a=(x@c.t()).argmax(1)
    return torch.argsort(a, stable=True)
TN, TU, S = 128, 256, 16
for mode in ('random','kmeans'):
    if mode=='kmeans':
        io = kmeans_order(hi, max(4, I//TN//2)); uo = kmeans_order(hu, max(4, U//TU//2))
    else:
        io = torch.arange(I); uo=torch.arange(U)
    hi2, hu2 = hi[io], hu[uo]
    nt = (I+TN-1)//TN
    cents = torch.stack([hi2[t*TN:(t+1)*TN].mean(0) for t in range(nt)])
    rad = torch.stack([(hi2[t*TN:(t+1)*TN]-cents[t]).norm(dim=1).max() for t in range(nt)])
    visited=0; total=0
    ng = min((U+TU-1)//TU, 40)
    for gi in range(ng):
        xu = hu2[gi*TU:(gi+1)*TU]
        ub = xu@cents.t() + rad[None,:]            # [TU, nt] upper bound of any score in tile
        order = torch.argsort(-ub.max(0).values)    # visit tiles by decreasing max bound
        tau = torch.full((xu.shape[0],), -1e9)
        top = torch.full((xu.shape[0], S), -1e9)
        for t in order.tolist():
            need = (ub[:,t] >= tau)                 # users for which the tile cannot be skipped
            total+=1
            if not need.any(): continue
            visited+=1
            sc = xu@hi2[t*TN:(t+1)*TN].t()
            top = torch.topk(torch.cat([top, sc],1), S, dim=1).values
            tau = top[:,-1]
    print('%s: U=%d I=%d tiles=%d  visited fraction %.3f  (mean tile radius %.3f)'%(mode,U,I,nt,visited/total, rad.mean()))
