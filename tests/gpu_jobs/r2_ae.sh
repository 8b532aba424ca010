#!/bin/bash
# round 2, call AE (1 GPU): Speech2Gesture baseline on hardware
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
echo "== pytest s2g"; timeout -s KILL 900 python -m pytest tests/test_gpu_zzzz_speech2gesture.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2ae_pytest_s2g.log 2>&1; echo "rc=$?"; grep -E "speech2gesture |passed|failed|^E " gpurun_out/r2ae_pytest_s2g.log | head -30 | cut -c1-300
echo "== step time"; timeout -s KILL 300 python - <<'PY' 2>&1 | grep -v Warn | tail -5
import argparse, os, sys, time
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from model.speech2gesture import Generator, Discriminator
from train_eval.train_speech2gesture import train_iter_speech2gesture
from tgb200 import config
dev = torch.device('cuda:0')
for mode in ('tf32', 'fp32'):
    config.set_mode(mode)
    torch.manual_seed(0)
    G, D = Generator(34, 27, 4).to(dev), Discriminator(27).to(dev)
    G.train(); D.train()
    args = argparse.Namespace(n_pre_poses=4, loss_regression_weight=100.0, loss_gan_weight=10.0)
    go = torch.optim.Adam(G.parameters(), lr=1e-3, betas=(0.5, 0.999)); do = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    B = 128
    spec = torch.randn(B, 128, 70, device=dev) * 20 - 40; tgt = torch.randn(B, 34, 27, device=dev) * 0.3
    for i in range(3):
        r = train_iter_speech2gesture(args, spec, tgt, G, D, go, do, None)
    torch.cuda.synchronize(); t0 = time.time()
    n = 5
    for i in range(n):
        r = train_iter_speech2gesture(args, spec, tgt, G, D, go, do, None)
    torch.cuda.synchronize(); dt = (time.time() - t0) / n
    print('speech2gesture batch 128 %s: %.2f ms/step = %.0f samples/s' % (mode, dt * 1e3, B / dt), r, 'max mem %.1f GB' % (torch.cuda.max_memory_allocated() / 2**30))
PY
