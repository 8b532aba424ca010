#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 600 python - <<'PY' 2>&1 | grep -v Warn | tail -30
import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from tgb200 import config, ops
from tgb200.engine import mm_nn, mm_nt, wgrad
dev = torch.device('cuda:0')
config.set_mode('fp32')
rel = lambda a, b: ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
g = torch.Generator().manual_seed(0)
B, W, C, N, k = 8, 34, 256, 256, 3
M, K = B * W, k * C
d = torch.randn(M, N, generator=g).to(dev)
w2 = (torch.randn(N, K, generator=g) * 0.05).to(dev)
dcol = torch.full((M, K), float('nan'), device=dev)
mm_nn(d, w2, None, dcol, M=M, N=N, K=K)
print('dgrad GEMM [272x256]x[256x768] fp32 vs fp64 matmul:', rel(dcol, d.double() @ w2.double()))
dx = torch.full((B, 1, W, C), float('nan'), device=dev)
ops.col2im2d(dcol, dx, B, 1, W, C, 1, k, 1, 1, 0, 1, 1, W)
ref = torch.zeros(B, W, C, dtype=torch.float64, device=dev)
dc = dcol.double().view(B, W, k, C)
for j in range(k):
    for wo in range(W):
        w = wo + j - 1
        if 0 <= w < W:
            ref[:, w] += dc[:, wo, j]
print('col2im vs loop:', rel(dx.view(B, W, C), ref))
# the forward GEMM and the weight gradient at the same shape
col = torch.randn(M, K, generator=g).to(dev)
y = torch.empty(M, N, device=dev)
mm_nt(col, w2, y, M=M, N=N, K=K)
print('forward GEMM:', rel(y, col.double() @ w2.double().t()))
dw = torch.zeros(N, K, device=dev)
wgrad(col, d, dw, B=B, T=W, N=N, Cin=K)
print('wgrad:', rel(dw, d.double().t() @ col.double()))
PY
