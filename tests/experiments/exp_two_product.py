"""Experiment (CPU, oracle embeddings): would a 2-product scoring scheme be sound often enough?

Scheme under test: score ~= (x_hi + x_lo) . yq, yq = round_to_16bit(y - c)  (user row exact to ~22 bits, item row
rounded ONCE).  Error of the approximate score: e(u,i) = x_u . ((y_i - c) - yq_i), |e| <= |x_u| * rho_i,
rho_i = ||(y_i - c) - yq_i||_2 (known exactly at prep time).  A user is SOUND for a shortlist of S when the k-th
exact score among the S best approximate ones is >= tau_S + rho_max - tie_tol, tau_S = S-th best approximate score.
Prints the fraction of users that would overflow for fp16 / bf16 item rounding and several S.

usage: python tests/experiments/exp_two_product.py U I E [n_users]
"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gnn_recsys_b200 as grb
from oracle import straightline as O
torch.manual_seed(1)
U, I, E = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
d = grb.make_graph(U, I, E, 0)
rel = d.relations()
num = {'user': U, 'item': I}
blk = O.block_from_coo(num, num, {c: (s.astype(np.int64), t.astype(np.int64), None) for c, (s, t) in rel.items()})
D = 128
sd = {}
for t, f in (('user', 2), ('item', 4)):
    l = torch.nn.Linear(f, D)
    sd['%s_embed.proj_feats.weight' % t] = l.weight.detach(); sd['%s_embed.proj_feats.bias' % t] = l.bias.detach()
gain = torch.nn.init.calculate_gain('relu')
for et in ('buys', 'bought-by', 'clicks', 'clicked-by'):
    for nm in ('fc_self', 'fc_neigh'):
        w = torch.empty(D, D); torch.nn.init.xavier_uniform_(w, gain=gain); sd['layers.0.mods.%s.%s.weight' % (et, nm)] = w
feats = {'user': d.user_feat, 'item': d.item_feat}
y = O.get_embeddings_full(num, [blk], feats, sd, D)
hu = torch.nn.functional.normalize(y['user'], dim=1); hi = torch.nn.functional.normalize(y['item'], dim=1)
c = hi.mean(0)
yc = hi - c
g = torch.Generator().manual_seed(5)
users = torch.randperm(U, generator=g)[:n]
x = hu[users].double()
exact = (x @ yc.double().t())            # centred exact scores (ranking-equal to the cosine)
k = 10
tie = 1e-5
print('U %d I %d E %d users sampled %d; |y-c| max %.4f mean %.4f' % (U, I, E, n, yc.norm(dim=1).max(), yc.norm(dim=1).mean()))
for name, dt in (('fp16', torch.float16), ('bf16', torch.bfloat16)):
    yq = yc.to(dt).float()
    rho = (yc - yq).norm(dim=1)
    rho_max = float(rho.max())
    approx = x @ yq.double().t()
    err = (approx - exact).abs()
    print('%s: rho_max %.3e rho_mean %.3e ; observed max|err| %.3e' % (name, rho_max, float(rho.mean()), float(err.max())))
    for S in (16, 24, 32, 48, 64):
        av, ai = torch.topk(approx, S, dim=1)
        tau = av[:, S - 1]
        ex_short = torch.gather(exact, 1, ai)
        sk = torch.topk(ex_short, k, dim=1).values[:, k - 1]
        overflow = (sk < tau + rho_max - tie)
        # same with the 3-product bound of the shipped kernel for reference (err 1.1e-5 * |y-c|max)
        print('   S=%2d: overflow fraction %.4f' % (S, float(overflow.float().mean())))
# three-product reference (shipped): err_rel 3*2^-18
e3 = 3 * 2.0 ** -18 * float(yc.norm(dim=1).max())
for S in (16,):
    av, ai = torch.topk(exact, S, dim=1)
    sk = av[:, k - 1]; tau = av[:, S - 1]
    print('3-product bf16 (err %.2e) S=16: overflow fraction %.4f' % (e3, float((sk < tau + e3 - tie).float().mean())))
# ---- single product, fp16: both sides rounded once. |err| <= rho_u * max|yq| + |x_u| * rho_max (per-user bound)
xq = x.float().to(torch.float16).double()
rho_u = (x - xq).norm(dim=1)
yq = yc.to(torch.float16).float()
rho_i = float((yc - yq).norm(dim=1).max())
ymax = float(yq.norm(dim=1).max())
approx = xq @ yq.double().t()
bound = rho_u * ymax + x.norm(dim=1) * rho_i
print('1-product fp16: per-user bound mean %.3e max %.3e ; observed max|err| %.3e' % (float(bound.mean()), float(bound.max()), float((approx - exact).abs().max())))
for S in (16, 24, 32, 48, 64):
    av, ai = torch.topk(approx, S, dim=1)
    tau = av[:, S - 1]
    sk = torch.topk(torch.gather(exact, 1, ai), k, dim=1).values[:, k - 1]
    print('   S=%2d: overflow fraction %.4f' % (S, float((sk < tau + bound - tie).float().mean())))
