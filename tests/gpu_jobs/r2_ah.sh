#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 600 python - <<'PY' 2>&1 | grep -v Warn | tail -120
import argparse, os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
from oracle import s2g_oracle as SO, synth
from oracle.make_golden_s2g import *
from model.speech2gesture import Generator, Discriminator
from train_eval.train_speech2gesture import train_iter_speech2gesture
from tgb200 import config
dev = torch.device('cuda:0')
config.set_mode('fp32')
rel = lambda a, b: ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
G, Dn = Generator(T, D, N_PRE), Discriminator(D)
gsd, dsd = synth.s2g_state_dict(G.state_dict(), G_SEED), synth.s2g_state_dict(Dn.state_dict(), D_SEED)
G.load_state_dict(gsd); Dn.load_state_dict(dsd); G.to(dev).train(); Dn.to(dev).train()
Bb = 8
spec, target = make_inputs(Bb, 41)
f64 = lambda sd: {k: (v.to(dev).double() if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
want = SO.train_iter_oracle(f64(gsd), f64(dsd), {}, {}, 1, spec.to(dev).double(), target.to(dev).double(), N_PRE, W_REG, W_GAN, LR, LR * D_LR_W)
args = argparse.Namespace(n_pre_poses=N_PRE, loss_regression_weight=W_REG, loss_gan_weight=W_GAN)
go = torch.optim.Adam(G.parameters(), lr=LR, betas=(0.5, 0.999)); do = torch.optim.Adam(Dn.parameters(), lr=LR * D_LR_W, betas=(0.5, 0.999))
g_before = {k: v.clone() for k, v in G.state_dict().items()}
ret = train_iter_speech2gesture(args, spec.to(dev), target.to(dev), G, Dn, go, do, None)
print(ret, want['losses'])
# D' of the GPU run vs the oracle's D'
for k, v in Dn.state_dict().items():
    if v.is_floating_point():
        w = want['d_sd'][k]
        print('D post %-28s rel %.2e  max|diff| %.2e' % (k, rel(v, w), (v.double() - w).abs().max().item()))
# gradient of the generator step recomputed in fp64 WITH THE GPU's OWN D' and the GPU's own generator output
ws = G.engine().ws
out = ws['g.final.y'].view(Bb, T, D).double()
dprime = f64({k: v.cpu() for k, v in Dn.state_dict().items()})
# BN running stats of D' were updated by the g pass; train-mode forward does not read them
outr = out.clone().requires_grad_(True)
om = outr[:, 1:] - outr[:, :-1]
s = SO.discriminator_forward(dprime, om, True, {})
loss = W_REG * (outr - target.to(dev).double()).abs().mean() + W_GAN * F.mse_loss(torch.ones_like(s), s)
loss.backward()
print('d_out (GPU) vs fp64 autograd through the GPU D-prime:', rel(ws['s2g.d_out'], outr.grad))
# generator backward given that d_out: compare G grads with fp64 autograd of the oracle generator fed the same d_out
gp = {k: v.to(dev).double().clone().requires_grad_(True) for k, v in g_before.items() if v.is_floating_point() and 'running' not in k}
gfull = f64({k: v.cpu() for k, v in g_before.items()}); gfull.update(gp)
o2 = SO.generator_forward(gfull, spec.to(dev).double(), target[:, :N_PRE].to(dev).double(), T, True, {})
(o2 * outr.grad).sum().backward()
for k, p in reversed(list(G.named_parameters())):
    if gp[k].grad.norm() > 1e-6: print('%-46s %.2e' % (k, rel(p.grad, gp[k].grad)))
PY
