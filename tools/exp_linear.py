"""gr_linear_f32 on the fc_preagg shapes: tcgen05 path vs the legacy 3xTF32 mma.sync path, time and max error vs fp64.

    python tools/exp_linear.py [n_rows]
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnn_recsys_b200 as grb
from gnn_recsys_b200 import ops
dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_500_000
g = torch.Generator(device=dev).manual_seed(0)
for d in (128, 256):
    x = torch.randn(n, d, device=dev, generator=g) * torch.exp(2 * torch.randn(n, 1, device=dev, generator=g))
    w = torch.randn(d, d, device=dev, generator=g) * (2.0 / d) ** 0.5
    wt = w.t().contiguous()
    ref = torch.relu(x[:4096].double() @ wt.double())
    for legacy in (False, True):
        for _ in range(2):
            y = ops.linear(x, wt, None, True, legacy=legacy)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = ops.linear(x, wt, None, True, legacy=legacy)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        err = ((y[:4096].double() - ref).abs() / (ref.abs() + 1e-5 * x[:4096].double().abs().amax(1, keepdim=True))).max().item()
        print('d=%d n=%d %-8s %.3f ms  %.1f useful TFLOP/s  %.0f GB/s (8 D bytes/row)  max rel err %.2e'
              % (d, n, 'legacy' if legacy else 'tcgen05', ms, 2.0 * n * d * d / ms / 1e9, 8.0 * d * n / ms / 1e6, err))
