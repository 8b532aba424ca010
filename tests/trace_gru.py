"""Development aid (not a test): per-step phase timeline of the tensor-core GRU kernels at the benchmark shape."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
import torch
from tgb200 import _lib, ops
dev = torch.device('cuda:0')
lib = _lib.load()
def run(kind, B, T=34, H=300, I=600):
    M = B * T
    g = torch.Generator().manual_seed(0)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(dev)
    gi = r(M, 6 * H); whh = [r(3 * H, H), r(3 * H, H)]; bhh = [r(3 * H), r(3 * H)]
    out = torch.zeros(M, 2 * H, device=dev); saved = torch.zeros(4, M, 2 * H, device=dev)
    sync = torch.zeros(64, dtype=torch.int32, device=dev)
    trace = torch.zeros(16 * T, dtype=torch.int64, device=dev)
    for rep in range(3):
        if kind == 'fwd':
            lib.tg_debug_gru_trace(ctypes.c_void_p(trace.data_ptr() if rep == 2 else 0))
            ops.gru_layer_fwd_tf32(gi, whh[0], whh[1], bhh[0], bhh[1], out, saved, M * 2 * H, sync, B, T, H)
        else:
            ops.gru_layer_fwd_tf32(gi, whh[0], whh[1], bhh[0], bhh[1], out, saved, M * 2 * H, sync, B, T, H)
            dout = r(M, 2 * H); dgi = torch.empty(M, 6 * H, device=dev); dgh = torch.empty(M, 6 * H, device=dev)
            partial = torch.empty(ops.gru_bwd_tf32_scratch_floats(B, H), device=dev)
            whhT = [w.t().contiguous() for w in whh]
            lib.tg_debug_gru_trace(ctypes.c_void_p(trace.data_ptr() if rep == 2 else 0))
            ops.gru_layer_bwd_tf32(dout, out, saved[0], M * 2 * H, whhT[0], whhT[1], dgi, dgh, partial, sync, B, T, H)
        torch.cuda.synchronize()
    lib.tg_debug_gru_trace(ctypes.c_void_p(0))
    tr = trace.cpu().view(T, 16).numpy()
    print(kind, 'B', B)
    base = tr[1][0] if kind == 'fwd' else tr[0][0]
    for s in range(0, T):
        row = tr[s]
        if row.max() == 0: continue
        t0 = row[row > 0].min()
        print('step %2d start %+8.2f us |' % (s, (t0 - base) / 1e3), ' '.join('%d:%6.2f' % (i, (v - t0) / 1e3) for i, v in enumerate(row) if v > 0))
run('fwd', 384); run('bwd', 128)
