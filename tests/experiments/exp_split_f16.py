import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import gnn_recsys_b200 as grb
from oracle import straightline as O
torch.manual_seed(1)
U,I,E,D = 10000,5000,200000,128
d = grb.make_graph(U,I,E,0); rel = d.relations()
sd={}
for t,f in (('user',2),('item',4)):
    l=torch.nn.Linear(f,D); sd['%s_embed.proj_feats.weight'%t]=l.weight.detach(); sd['%s_embed.proj_feats.bias'%t]=l.bias.detach()
gain=torch.nn.init.calculate_gain('relu')
for et in ('buys','bought-by','clicks','clicked-by'):
    for nm in ('fc_self','fc_neigh'):
        w=torch.empty(D,D); torch.nn.init.xavier_uniform_(w,gain=gain); sd['layers.0.mods.%s.%s.weight'%(et,nm)]=w
h = O.embed_inputs({'user':d.user_feat,'item':d.item_feat}, sd)
def pow2scale(m):
    m = m.clamp(min=1e-30)
    return torch.exp2(-torch.floor(torch.log2(m)))
def split16(x):
    hi = x.half().float(); lo=(x-hi).half().float(); return hi, lo
worst=0
for scale_in in (1.0, 1e-4, 3e4):
  for c,(s,t) in rel.items():
    s=torch.from_numpy(s.astype(np.int64)); t=torch.from_numpy(t.astype(np.int64))
    hs, hd = h[c[0]]*scale_in, h[c[2]]*scale_in
    n = O.neighbour_reduce(s,t,None,hs,hd.shape[0],'mean')
    ws, wn = sd['layers.0.mods.%s.fc_self.weight'%c[1]], sd['layers.0.mods.%s.fc_neigh.weight'%c[1]]
    A = torch.cat([hd, n], 1); W = torch.cat([ws, wn], 1).t().contiguous()   # [256, 128]
    z64 = torch.relu(A.double()@W.double()); z64 = z64/ z64.norm(dim=1,keepdim=True).clamp(min=1e-300)
    sr = pow2scale(A.abs().max(1, keepdim=True).values); sw = pow2scale(W.abs().max())
    ah, al = split16(A*sr); wh, wl = split16(W*sw)
    z = (al.double()@wh.double() + ah.double()@wl.double() + ah.double()@wh.double()).float()
    z = torch.relu(z); z = z/ z.norm(dim=1,keepdim=True).clamp(min=1e-30)
    err = (z.double()-z64).abs(); tol = 1e-5 + 1e-4*z64.abs()
    worst=max(worst,float((err/tol).max()))
  print('input scale %g: max err/tol %.4f'%(scale_in, worst))
