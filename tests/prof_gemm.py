"""Profiling driver (not a test): the two dominant tcgen05 GEMM shapes in isolation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
import torch
from tgb200 import ops
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(dev)
M = 13056
a = r(M, 600); w = r(1800, 600); b = r(1800); c = torch.empty(M, 1800, device=dev)
x = r(M, 300); wt = r(600, 300); b2 = r(300); y = torch.empty(M, 300, device=dev); mask = torch.ones(M, 300, device=dev)
G = r(4352, 1800); X = r(4352, 600); dW = torch.zeros(1800, 600, device=dev)
for i in range(3):
    ops.gemm_tf32(a, w, c, M=M, N=1800, K=600, bias=b)
    ops.gemm_tf32(x, wt, y, M=M, N=300, K=300, taps=2, shift0=-2, T=34, bias=b2, act1=1, mask=mask)
    ops.wgrad_tf32(G, X, dW, B=128, T=34, N=1800, Cin=600)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, fl in (('ih', lambda: ops.gemm_tf32(a, w, c, M=M, N=1800, K=600, bias=b), 2.0 * M * 1800 * 600),
                     ('tcn', lambda: ops.gemm_tf32(x, wt, y, M=M, N=300, K=300, taps=2, shift0=-2, T=34, bias=b2, act1=1, mask=mask), 4.0 * M * 300 * 300),
                     ('wgrad', lambda: ops.wgrad_tf32(G, X, dW, B=128, T=34, N=1800, Cin=600), 2.0 * 4352 * 1800 * 600)):
    e0.record()
    for i in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('%s %.1f us  %.1f TFLOP/s' % (name, ms * 1e3, fl / ms / 1e9))
