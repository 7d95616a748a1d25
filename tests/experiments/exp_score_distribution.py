import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import gnn_recsys_b200 as grb
from oracle import straightline as O
torch.manual_seed(1)
U,I,E = int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
d = grb.make_graph(U,I,E,0)
rel = d.relations()
num={'user':U,'item':I}
blk = O.block_from_coo(num,num,{c:(s.astype(np.int64),t.astype(np.int64),None) for c,(s,t) in rel.items()})
D=128
sd={}
import math
def lin(o,i,bias=True):
    l=torch.nn.Linear(i,o,bias=bias); return l
for t,f in (('user',2),('item',4)):
    l=lin(D,f); sd['%s_embed.proj_feats.weight'%t]=l.weight.detach(); sd['%s_embed.proj_feats.bias'%t]=l.bias.detach()
gain=torch.nn.init.calculate_gain('relu')
for et in ('buys','bought-by','clicks','clicked-by'):
    for nm in ('fc_self','fc_neigh'):
        w=torch.empty(D,D); torch.nn.init.xavier_uniform_(w,gain=gain); sd['layers.0.mods.%s.%s.weight'%(et,nm)]=w
feats={'user':d.user_feat,'item':d.item_feat}
y=O.get_embeddings_full(num,[blk],feats,sd,D)
hu=torch.nn.functional.normalize(y['user'],dim=1); hi=torch.nn.functional.normalize(y['item'],dim=1)
c=hi.mean(0)
print('item norm after centering: max %.4f mean %.4f'%((hi-c).norm(dim=1).max(), (hi-c).norm(dim=1).mean()))
n=min(U,2000)
S=(hu[:n]@hi.t())
v,_=torch.topk(S,64,dim=1)
print('score top1 mean %.4f  median score %.4f'%(v[:,0].mean(), S.median()))
for r in (16,32,64):
    gap=(v[:,9]-v[:,r-1])
    print('gap rank10-rank%d: median %.2e  p10 %.2e  p90 %.2e  frac<1e-5 %.3f frac<1e-3 %.3f frac <8e-3 %.3f'%(r,gap.median(),gap.quantile(0.1),gap.quantile(0.9),(gap<1e-5).float().mean(),(gap<1e-3).float().mean(),(gap<8e-3).float().mean()))
# count of items within eps of the 10th score
for eps in (1e-5,1e-4,1e-3,4e-3):
    cnt=(S>=(v[:,9:10]-2*eps)).sum(1).float()
    print('eps %.0e: #items within 2eps of 10th: median %d p90 %d max %d'%(eps,cnt.median(),cnt.quantile(0.9),cnt.max()))
print('unique item rows: %d of %d'%(torch.unique((hi*1e6).round(),dim=0).shape[0], I))
