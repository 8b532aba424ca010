"""Profiling driver (not a test): latency of the small / skinny GEMM shapes of the discriminator and the heads, tensor-core (TF32)
kernel vs. the fp32 FFMA kernel, each timed alone with CUDA events (50 back-to-back launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
import torch
from tgb200 import ops
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, n=50):
    for i in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (M, N, K) in [(3584, 384, 128), (3584, 128, 384), (3584, 384, 8), (3584, 8, 384), (13056, 1800, 600), (4352, 600, 1800), (13056, 1800, 108),
                  (13056, 32, 300), (4352, 300, 32), (128, 384, 128)]:
    a = r(M, K); w = r(N, K); b = r(N); c = torch.empty(M, N, device=dev)
    t_tc = timeit(lambda: ops.gemm_tf32(a, w, c, M=M, N=N, K=K, bias=b))
    t_f = timeit(lambda: ops.conv_gemm(a, w, c, B=1, Tin=M, Tout=M, N=N, Cin=K, taps=1, lda=K, ldw=K, wsc=1, bias=b))
    fl = 2.0 * M * N * K
    print('M%-6d N%-5d K%-5d  tf32 %7.1f us (%6.1f TF/s)   fp32 %7.1f us (%6.1f TF/s)' % (M, N, K, t_tc, fl / t_tc / 1e6, t_f, fl / t_f / 1e6))
x = r(13056, 300); wt = r(600, 300); b2 = r(300); y = torch.empty(13056, 300, device=dev)
t = timeit(lambda: ops.gemm_tf32(x, wt, y, M=13056, N=300, K=300, taps=2, shift0=-2, T=34, bias=b2, act1=1))
print('tcn 2-tap M13056 N300 K300: %.1f us (%.1f TF/s)' % (t, 4.0 * 13056 * 300 * 300 / t / 1e6))
for (B, T, N, Cin, shift) in [(128, 28, 192, 64, 1), (128, 28, 384, 128, 0), (128, 28, 384, 8, 0), (128, 34, 1800, 600, 0), (128, 34, 900, 300, -1),
                              (128, 34, 300, 300, -2)]:
    G = r(B * T, N); X = r(B * T, Cin); dW = torch.zeros(N, Cin, device=dev)
    t = timeit(lambda: ops.wgrad_tf32(G, X, dW, B=B, T=T, N=N, Cin=Cin, shift=shift))
    print('wgrad B%d T%d N%d Cin%d shift%d: %.1f us (%.1f TF/s)' % (B, T, N, Cin, shift, t, 2.0 * B * T * N * Cin / t / 1e6))
