"""Profiling driver (not a test): ONE launch of each dominant tensor-core kernel shape of the G+D step, for `ncu --set full`.
Order: ih GEMM (GRU input projection), TCN two-tap GEMM, WavEncoder conv2 window GEMM, conv2 column GEMM (data gradient),
weight gradients: GRU weight_ih [1800x600], TCN tap, WavEncoder conv2 window view."""
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
B, Tin, cin, cout, k, s = 128, 7891, 16, 32, 15, 6
Tout = (Tin - k) // s + 1
act = r(B * Tin, cin); w2 = r(cout, k * cin); bb = r(cout); y2 = torch.empty(B * Tout, cout, device=dev)
dy2 = r(B * Tout, cout); w2t = r(k * cin, cout); col = torch.empty(B * Tout, k * cin, device=dev)
G1 = r(4352, 1800); X1 = r(4352, 600); dW1 = torch.zeros(1800, 600, device=dev)
G2 = r(4352, 300); X2 = r(4352, 300); dW2 = torch.zeros(300, 300, device=dev)
dW3 = torch.zeros(cout, k * cin, device=dev)
torch.cuda.synchronize()
ops.gemm_tf32(a, w, c, M=M, N=1800, K=600, bias=b)
ops.gemm_tf32(x, wt, y, M=M, N=300, K=300, taps=2, shift0=-2, T=34, bias=b2, act1=1, mask=mask)
ops.gemm_tf32(act, w2, y2, M=B * Tout, N=cout, K=k * cin, lda=s * cin, clip_rows=Tout, a_clip_pitch=Tin * cin, bias=bb)
ops.gemm_tf32(dy2, w2t, col, M=B * Tout, N=k * cin, K=cout)
ops.wgrad_tf32(G1, X1, dW1, B=128, T=34, N=1800, Cin=600)
ops.wgrad_tf32(G2, X2, dW2, B=128, T=34, N=300, Cin=300, shift=-2)
ops.wgrad_tf32(dy2, act, dW3, B=B, T=Tout, N=cout, Cin=k * cin, ldx=s * cin, x_clip_pitch=Tin * cin)
torch.cuda.synchronize()
print('done')
