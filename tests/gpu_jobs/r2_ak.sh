#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout -s KILL 600 python - <<'PY' 2>&1 | grep -v Warn | tail -30
import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
from oracle import s2g_oracle as SO, synth
from oracle.make_golden_s2g import *
from model.speech2gesture import Generator
dev = torch.device('cuda:0')
G = Generator(T, D, N_PRE)
gsd = synth.s2g_state_dict(G.state_dict(), G_SEED)
spec, target = make_inputs(8, 41)
g = torch.Generator().manual_seed(1); dout = (torch.randn(8, T, D, generator=g) * 0.01).to(dev)
res = {}
for dt in (torch.float32, torch.float64):
    sd = {k: (v.to(dev).to(dt) if v.is_floating_point() else v.to(dev)) for k, v in gsd.items()}
    gp = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and 'running' not in k}
    full = dict(sd); full.update(gp)
    o = SO.generator_forward(full, spec.to(dev).to(dt), target[:, :N_PRE].to(dev).to(dt), T, True, {})
    (o * dout.to(dt)).sum().backward()
    res[dt] = {k: v.grad for k, v in gp.items()}
print('stock PyTorch on the same GPU, fp32 (TF32 off) vs fp64, same inputs:')
for k in ('final_out.weight', 'decoder.3.0.weight', 'decoder.2.1.bias', 'decoder.2.1.weight', 'decoder.2.0.weight', 'decoder.0.0.weight', 'pre_pose_encoder.0.weight', 'audio_encoder.up5.conv.0.weight', 'audio_encoder.down6.1.bias', 'audio_encoder.first_net.0.0.weight'):
    a, b = res[torch.float32][k].double(), res[torch.float64][k]
    print('  %-40s %.2e' % (k, ((a - b).norm() / b.norm()).item()))
PY
