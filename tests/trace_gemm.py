"""Development aid (not a test): phase timeline of CTA (0,0) of the tcgen05 GEMM at small shapes (where does a 20 us kernel spend its time?)."""
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
for (M, N, K) in [(3584, 384, 128), (3584, 128, 384), (128, 128, 32), (13056, 1800, 600), (13056, 300, 300)]:
    a = r(M, K); w = r(N, K); b = r(N); c = torch.empty(M, N, device=dev)
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    for rep in range(4):
        lib.tg_debug_gemm_trace(ctypes.c_void_p(trace.data_ptr() if rep == 3 else 0))
        torch.cuda.synchronize()
        e0.record()
        ops.gemm_tf32(a, w, c, M=M, N=N, K=K, bias=b)
        e1.record()
        torch.cuda.synchronize()
    lib.tg_debug_gemm_trace(ctypes.c_void_p(0))
    t = trace.cpu().numpy()
    print('M%d N%d K%d: event %.1f us | cta(0,0) stamps (us from kernel start): ' % (M, N, K, e0.elapsed_time(e1) * 1e3) +
          ' '.join('%d:%.2f' % (i, (t[i] - t[0]) / 1e3) for i in range(7)))
