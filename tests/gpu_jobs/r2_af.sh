#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 600 python - <<'PY' 2>&1 | grep -v Warn | tail -80
import argparse, os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from oracle import s2g_oracle as SO, synth
from oracle.make_golden_s2g import *
from model.speech2gesture import Generator, Discriminator
from train_eval.train_speech2gesture import train_iter_speech2gesture
from tgb200 import config
dev = torch.device('cuda:0')
config.set_mode('fp32')
G, Dn = Generator(T, D, N_PRE), Discriminator(D)
gsd, dsd = synth.s2g_state_dict(G.state_dict(), G_SEED), synth.s2g_state_dict(Dn.state_dict(), D_SEED)
G.load_state_dict(gsd); Dn.load_state_dict(dsd); G.to(dev).train(); Dn.to(dev).train()
Bb = 8
spec, target = make_inputs(Bb, 41)
f64 = lambda sd: {k: (v.to(dev).double() if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
want = SO.train_iter_oracle(f64(gsd), f64(dsd), {}, {}, 1, spec.to(dev).double(), target.to(dev).double(), N_PRE, W_REG, W_GAN, LR, LR * D_LR_W)
args = argparse.Namespace(n_pre_poses=N_PRE, loss_regression_weight=W_REG, loss_gan_weight=W_GAN)
go = torch.optim.Adam(G.parameters(), lr=LR, betas=(0.5, 0.999)); do = torch.optim.Adam(Dn.parameters(), lr=LR * D_LR_W, betas=(0.5, 0.999))
ret = train_iter_speech2gesture(args, spec.to(dev), target.to(dev), G, Dn, go, do, None)
print(ret, want['losses'])
for k, p in G.named_parameters():
    r = want['g_grads'][k]
    e = ((p.grad.double() - r).norm() / (r.norm() + 1e-12)).item()
    print('%-50s |ref| %.3e  rel %.2e' % (k, r.norm().item(), e))
PY
