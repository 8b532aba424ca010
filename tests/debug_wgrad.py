import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
import torch
from tgb200 import ops
dev = torch.device('cuda:0')
torch.manual_seed(0)
for (T, N, C) in [(32, 128, 128), (8, 128, 128), (32, 32, 32)]:
    G = torch.randint(-3, 4, (T, N), device=dev).float()
    X = torch.randint(-3, 4, (T, C), device=dev).float()
    dW = torch.zeros(N, C, device=dev)
    ops.wgrad_tf32(G, X, dW, B=1, T=T, N=N, Cin=C)
    torch.cuda.synchronize()
    ref = G.t() @ X
    print('T,N,C', T, N, C, 'max|dW|', dW.abs().max().item(), 'max|ref|', ref.abs().max().item(), 'err', (dW - ref).abs().max().item())
    print(' dW[0,:8]', dW[0, :8].tolist())
    print(' ref[0,:8]', ref[0, :8].tolist())
    print(' dW[:8,0]', dW[:8, 0].tolist())
    # hypotheses
    print(' err vs ref^T', (dW - ref.t()).abs().max().item() if N == C else None)
    nz = (dW != 0).float().mean().item()
    print(' nonzero frac', nz)
    # single-k contribution probes: which (k) rows are used
    for k in range(min(T, 9)):
        Gk = torch.zeros_like(G); Gk[k] = G[k]
        d2 = torch.zeros(N, C, device=dev)
        ops.wgrad_tf32(Gk, X, d2, B=1, T=T, N=N, Cin=C)
        torch.cuda.synchronize()
        r2 = Gk.t() @ X
        print('  k=%d max|d2| %.1f err %.1f' % (k, d2.abs().max().item(), (d2 - r2).abs().max().item()))
