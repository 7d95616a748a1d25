import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import gnn_recsys_b200 as grb
from gnn_recsys_b200 import ops, _native as N
dev = torch.device('cuda:0')
for n, k, m in ((2_500_000, 256, 256), (1_000_000, 128, 128)):
    x = torch.randn(n, k, device=dev); w = torch.randn(k, m, device=dev) * 0.1
    def run(tc):
        y = torch.empty(n, m, device=dev)
        ws = ops._ws(N.load().gr_linear_workspace_bytes(k, m), dev, 'linear')
        N.call('gr_linear_f32', N.ptr(x), n, k, N.ptr(w), None, m, 1, N.ptr(y), N.ptr(ws) if tc else None, ws.numel() if tc else 0, N.stream())
        return y
    for tc in (False, True):
        for _ in range(2): y = run(tc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): y = run(tc)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ref = torch.relu(x[:4096].double() @ w.double())
        err = (y[:4096].double() - ref).abs().max().item()
        print('n=%d k=%d m=%d %s: %.2f ms  %.1f TFLOP/s  max abs err %.2e' % (n, k, m, '3xTF32' if tc else 'FFMA', ms, 2.0*n*k*m/ms/1e9, err))
