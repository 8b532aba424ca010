"""Profiling driver (not a test): where does the end-to-end step (inputs from pinned host memory through DevicePrefetcher) lose time
against the resident-input step?  Prints, per step: first/last kernel time, the copy kernel's interval, and idle gaps > 30 us."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from model import vocab  # noqa: E402
from model.multimodal_context_net import ConvDiscriminator, PoseGenerator  # noqa: E402
from train_eval.staging import DevicePrefetcher  # noqa: E402
from train_eval.train_gan import train_iter_gan  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
args = bench.make_args_ns()
spk = vocab.Vocab('vid', insert_default_tokens=False)
while spk.n_words < bench.N_SPEAKERS:
    spk.index_word('s%d' % spk.n_words)
G = PoseGenerator(args, bench.POSE_DIM, bench.N_WORDS, 300, None, z_obj=spk).to(dev).train()
D = ConvDiscriminator(bench.POSE_DIM).to(dev).train()
g_opt = torch.optim.Adam(G.parameters(), lr=5e-4, betas=(0.5, 0.999))
d_opt = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.999))
pinned = [{k: v.pin_memory() for k, v in bench.synth_batch(128, i).items()} for i in range(4)]


def run(n, prof=None):
    it = iter(DevicePrefetcher((pinned[i % 4] for i in range(n)), dev))
    t = []
    for i in range(n):
        t0 = time.perf_counter()
        b = next(it)
        t1 = time.perf_counter()
        train_iter_gan(args, 11, b['in_text'], b['in_audio'], b['target'], b['vid'], G, D, g_opt, d_opt)
        t2 = time.perf_counter()
        t.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    return t


run(8)
torch.cuda.synchronize()
print('host ms per step (prefetcher next, train_iter_gan incl. sync):', [(round(a, 3), round(b, 3)) for a, b in run(6)])
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(4)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
last_end = None
for e in evs:
    s, en = e.time_range.start - t0, e.time_range.end - t0
    if 'copy_bytes' in e.name or 'Memcpy HtoD' in e.name:
        print('%9.1f %8.1f  COPY %s' % (s, en - s, e.name[:50]))
    if last_end is not None and s - last_end > 30:
        print('%9.1f  gap %.1f us before %s' % (s, s - last_end, e.name[:60]))
    last_end = en if last_end is None else max(last_end, en)
print('span per step us', (evs[-1].time_range.end - t0) / 4)
