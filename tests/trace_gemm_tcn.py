"""Development aid (not a test): phase timeline of CTA (0,0) of the two-tap TCN GEMM (csrc/gemm_tcn.cu) and, with TGB200_TCN_TWO_ACC=1, of
the two-accumulator tiles it replaces.  Stamps: 0 start, 1 setup done, 2 first stage landed, 7 first stage of tap 1, 3 last MMA issued,
4 accumulator ready, 5 epilogue done, 6 exit."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
import torch
from tgb200 import _lib, ops
dev = torch.device('cuda:0')
lib = _lib.load()
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
T = 34
for (B, N, K, epi) in [(384, 300, 300, 'bias'), (384, 300, 300, 'mask+res'), (128, 300, 300, 'bias')]:
    M = B * T
    a = r(M, K); w = r(2 * N, K); b = r(N); c = torch.empty(M, N, device=dev); mk = r(M, N); res = r(M, N)
    kw = dict(bias=b, act1=ops.ACT_RELU)
    if epi != 'bias':
        kw.update(mask=mk, residual=res, act2=ops.ACT_RELU)
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    ts = []
    for rep in range(6):
        lib.tg_debug_gemm_trace(ctypes.c_void_p(trace.data_ptr() if rep == 5 else 0))
        torch.cuda.synchronize()
        e0.record()
        ops.gemm_tf32(a, w, c, M=M, N=N, K=K, taps=2, shift0=-2, T=T, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    lib.tg_debug_gemm_trace(ctypes.c_void_p(0))
    # back-to-back launches: GPU time per launch without the host's descriptor encoding in between
    torch.cuda.synchronize(); e0.record()
    for rep in range(20):
        ops.gemm_tf32(a, w, c, M=M, N=N, K=K, taps=2, shift0=-2, T=T, **kw)
    e1.record(); torch.cuda.synchronize()
    t = trace.cpu().numpy()
    print('B%d N%d K%d %s: single-launch event %.1f us, 20 back-to-back %.1f us each | cta(0,0) stamps (us): ' % (B, N, K, epi, min(ts), e0.elapsed_time(e1) * 50) +
          ' '.join('%d:%.2f' % (i, (t[i] - t[0]) / 1e3) for i in (1, 2, 7, 3, 4, 5, 6)))
