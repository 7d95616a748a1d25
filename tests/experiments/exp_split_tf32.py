import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import gnn_recsys_b200 as grb
from oracle import straightline as O
torch.manual_seed(1)
U,I,E,D = 10000,5000,200000,128
d = grb.make_graph(U,I,E,0); rel = d.relations(); num={'user':U,'item':I}
sd={}
for t,f in (('user',2),('item',4)):
    l=torch.nn.Linear(f,D); sd['%s_embed.proj_feats.weight'%t]=l.weight.detach(); sd['%s_embed.proj_feats.bias'%t]=l.bias.detach()
gain=torch.nn.init.calculate_gain('relu')
for et in ('buys','bought-by','clicks','clicked-by'):
    for nm in ('fc_self','fc_neigh'):
        w=torch.empty(D,D); torch.nn.init.xavier_uniform_(w,gain=gain); sd['layers.0.mods.%s.%s.weight'%(et,nm)]=w
h = O.embed_inputs({'user':d.user_feat,'item':d.item_feat}, sd)
def tf32(x):
    i = x.view(torch.int32); r = ((i + 0x1000) & ~0x1fff)  # round-to-nearest (ties away) to 10-bit mantissa
    return r.view(torch.float32)
def split(x, kind):
    if kind=='bf16': hi = x.to(torch.bfloat16).float(); lo=(x-hi).to(torch.bfloat16).float()
    elif kind=='fp16': hi = x.half().float(); lo=(x-hi).half().float()
    else: hi = tf32(x.contiguous()); lo = tf32((x-hi).contiguous())
    return hi, lo
def mm3(a, b, kind):
    ah,al = split(a,kind); bh,bl = split(b,kind)
    return (ah.double()@bh.double() + al.double()@bh.double() + ah.double()@bl.double()).float()
for kind in ('bf16','fp16','tf32'):
    worst=0; worst_abs=0
    for c,(s,t) in rel.items():
        s=torch.from_numpy(s.astype(np.int64)); t=torch.from_numpy(t.astype(np.int64))
        hs, hd = h[c[0]], h[c[2]]
        n = O.neighbour_reduce(s,t,None,hs,hd.shape[0],'mean')
        ws, wn = sd['layers.0.mods.%s.fc_self.weight'%c[1]], sd['layers.0.mods.%s.fc_neigh.weight'%c[1]]
        z64 = torch.relu(hd.double()@ws.double().t() + n.double()@wn.double().t()); z64 = z64/ z64.norm(dim=1,keepdim=True).clamp(min=1e-30)
        z = torch.relu(mm3(hd, ws.t().contiguous(), kind) + mm3(n, wn.t().contiguous(), kind)); z = z/ z.norm(dim=1,keepdim=True).clamp(min=1e-30)
        zf = torch.relu(hd@ws.t() + n@wn.t()); zf = zf/zf.norm(dim=1,keepdim=True).clamp(min=1e-30)
        err = (z.double()-z64).abs(); tol = 1e-5 + 1e-4*z64.abs()
        errf = (zf.double()-z64).abs()
        worst=max(worst, float((err/tol).max())); worst_abs=max(worst_abs,float(err.max()))
        wf = float((errf/tol).max())
    print(kind, 'max err/tol %.3f  max abs err %.2e   (plain fp32 err/tol %.3f)'%(worst, worst_abs, wf))
