"""Profiling driver (not a test): per-kernel device time of graph-replayed train_iter_seq2seq steps (CUPTI via torch.profiler)."""
import argparse, collections, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile
from model.seq2seq_net import Seq2SeqNet
from train_eval.train_seq2seq import train_iter_seq2seq
dev = torch.device('cuda:0')
a = argparse.Namespace(hidden_size=200, n_layers=2, dropout_prob=0.1, n_pre_poses=4, GAN_noise_size=0, loss_regression_weight=250.0,
                       loss_kld_weight=0.1, loss_reg_weight=25.0)
net = Seq2SeqNet(a, 27, 34, 20000, 300, None).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
rng = np.random.Generator(np.random.PCG64(7))
lengths = np.sort(rng.integers(4, 13, size=128))[::-1].copy(); lengths[0] = 12
text = np.zeros((128, 12), dtype=np.int64)
for b in range(128):
    n = int(lengths[b]); text[b, 0] = 1; text[b, 1:n - 1] = rng.integers(4, 20000, size=n - 2); text[b, n - 1] = 2
text = torch.from_numpy(text).to(dev); lens = torch.from_numpy(lengths.astype(np.int64))
target = (0.1 * torch.randn(128, 34, 27)).to(dev)
f = lambda: train_iter_seq2seq(a, 0, text, lens, target, net, opt)
for i in range(6):
    f()
print('graph captured:', any(s.graph is not None for s in net.engine()._graph_slots.values()))
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10):
    f()
torch.cuda.synchronize(); print('ms/step wall', (time.perf_counter() - t0) * 100)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        f()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
for e in evs:
    k = e.name.replace('(anonymous namespace)::', '')[:60]
    agg[k][0] += 1; agg[k][1] += e.time_range.end - e.time_range.start
tot = sum(v[1] for v in agg.values())
span = max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)
print('kernels/step %d, busy us/step %.0f, span us/step %.0f' % (len(evs) // 2, tot / 2, span / 2))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print('%-62s n/step %5d  us/step %8.1f  avg %6.1f' % (k, v[0] // 2, v[1] / 2, v[1] / v[0]))
