#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 600 python - <<'PY' 2>&1 | grep -v Warn | tail -60
import argparse, os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from oracle import s2g_oracle as SO, synth
from oracle.make_golden_s2g import *
from model.speech2gesture import Generator, Discriminator
from tgb200 import config
dev = torch.device('cuda:0')
config.set_mode('fp32')
rel = lambda a, b: ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
Dn = Discriminator(D)
dsd = synth.s2g_state_dict(Dn.state_dict(), D_SEED)
Dn.load_state_dict(dsd); Dn.to(dev).train()
for Bb in (8, 32):
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(Bb, 33, D, generator=g) * 0.3).to(dev)
    de = Dn.engine().ensure(dev)
    de.arena.zero_grad()
    s = de.forward(x, True, slot='t')
    n = s.numel()
    gs = (torch.randn(n, generator=g) * 0.1).to(dev)
    dp = de.backward(gs.clone().view(n, 1), need_dposes=True)
    f64 = {k: (v.to(dev).double() if v.is_floating_point() else v.to(dev)) for k, v in dsd.items()}
    params = {k: v.clone().requires_grad_(True) for k, v in f64.items() if v.is_floating_point() and 'running' not in k}
    full = dict(f64); full.update(params)
    xr = x.double().requires_grad_(True)
    out = SO.discriminator_forward(full, xr, True, {})
    print('B', Bb, 'scores', rel(s.view(Bb, -1), out.view(Bb, -1)))
    (out.view(-1) * gs.double()).sum().backward()
    print('   dposes', rel(dp, xr.grad))
    for k, p in Dn.named_parameters():
        print('   %-20s %.2e' % (k, rel(p.grad, params[k].grad)))
    # layer-by-layer activations
    h = (xr[:, 1:] - xr[:, :-1]).transpose(1, 2)
PY
