#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 600 python - <<'PY' 2>&1 | grep -v Warn | tail -60
import argparse, os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
from oracle import s2g_oracle as SO, synth
from oracle.make_golden_s2g import *
from model.speech2gesture import Generator
from tgb200 import config, s2g_engine
dev = torch.device('cuda:0')
config.set_mode('fp32')
rel = lambda a, b: ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
G = Generator(T, D, N_PRE)
gsd = synth.s2g_state_dict(G.state_dict(), G_SEED)
G.load_state_dict(gsd); G.to(dev).train()
Bb = 8
spec, target = make_inputs(Bb, 41)
ge = G.engine().ensure(dev)
rec = {}
orig_fwd, orig_bwd = s2g_engine._ConvBlock.forward, s2g_engine._ConvBlock.backward
def fwd(self, x, B, H, W, training):
    rec[self.tag + '.in'] = x.clone()
    out = orig_fwd(self, x, B, H, W, training)
    rec[self.tag + '.out'] = out[0].clone()
    return out
def bwd(self, d, need_dx=True, param_grads=True):
    rec[self.tag + '.dout'] = d.clone()
    dx = orig_bwd(self, d, need_dx, param_grads)
    if dx is not None: rec[self.tag + '.dx'] = dx.clone()
    return dx
s2g_engine._ConvBlock.forward, s2g_engine._ConvBlock.backward = fwd, bwd
out = ge.forward(spec.to(dev), target[:, :N_PRE].to(dev), True)
g = torch.Generator().manual_seed(1)
dout = (torch.randn(Bb, T, D, generator=g) * 0.01).to(dev)
ge.arena.zero_grad()
ge.backward(dout.clone())
torch.cuda.synchronize()
sd = {k: v.to(dev).double() if v.is_floating_point() else v.to(dev) for k, v in gsd.items()}
for blk in [ge.final] + list(reversed(ge.dec)) + list(reversed(ge.up)) + list(reversed(ge.down)) + list(reversed(ge.down1)) + list(reversed(ge.first))[:3]:
    Bq, H, W, cin, kh, kw, sh, sw, pt, pl, Ho, Wo, cout = blk.geom
    x = rec[blk.tag + '.in'].double().view(Bq, H, W, cin).permute(0, 3, 1, 2).clone().requires_grad_(True)
    w = sd[blk.conv + '.weight'].clone().requires_grad_(True); b = sd[blk.conv + '.bias']
    sdl = dict(sd)
    if blk.bn is not None:
        sdl[blk.bn + '.weight'] = sd[blk.bn + '.weight'].clone().requires_grad_(True); sdl[blk.bn + '.bias'] = sd[blk.bn + '.bias'].clone().requires_grad_(True)
    m = ge.mod(blk.conv)
    pad = 'VALID' if (getattr(m, 'padding', 0) == 'VALID' or getattr(m, 'padding', 0) in (0, (0,), (0, 0))) else 'SAME'
    xin = x if w.dim() == 4 else x.squeeze(2)
    y = SO.conv_tf(xin, w, b, m.stride, pad)
    if blk.bn is not None:
        y = F.leaky_relu(SO._bn(y, sdl, blk.bn, True, None), 0.2)
    yl = y if w.dim() == 4 else y.unsqueeze(2)
    fo = rel(rec[blk.tag + '.out'].view(Bq, Ho, Wo, cout).permute(0, 3, 1, 2), yl)
    dy = rec[blk.tag + '.dout'].double().view(Bq, Ho, Wo, cout).permute(0, 3, 1, 2)
    (yl * dy).sum().backward()
    line = '%-10s fwd %.1e' % (blk.tag, fo)
    if blk.tag + '.dx' in rec:
        line += '  dx %.2e' % rel(rec[blk.tag + '.dx'].view(Bq, H, W, cin).permute(0, 3, 1, 2), x.grad)
    pg = dict(G.named_parameters())
    line += '  dW %.2e' % rel(pg[blk.conv + '.weight'].grad, w.grad)
    if blk.bn is not None:
        line += '  dgamma %.2e dbeta %.2e' % (rel(pg[blk.bn + '.weight'].grad, sdl[blk.bn + '.weight'].grad), rel(pg[blk.bn + '.bias'].grad, sdl[blk.bn + '.bias'].grad))
    print(line)
gp = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and 'running' not in k}
full = dict(sd); full.update(gp)
o2 = SO.generator_forward(full, spec.to(dev).double(), target[:, :N_PRE].to(dev).double(), T, True, {})
(o2 * dout.double()).sum().backward()
for k, p in reversed(list(G.named_parameters())):
    if gp[k].grad.norm() > 1e-9 and ('decoder' in k or 'up5' in k or 'first_net.0' in k or 'pre_pose' in k): print('%-46s %.2e' % (k, rel(p.grad, gp[k].grad)))
PY
