"""Profiling driver (not a test): kernel timeline (start, duration, stream) of graph-replayed train_iter_gan steps,
taken with torch.profiler (CUPTI activity records; unlike ncu the kernels run concurrently as in the bench).
Writes gpurun_out/timeline.csv: step-relative start_us, dur_us, stream, kernel name."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (sets up sys.path for the package)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from model import vocab  # noqa: E402
from model.multimodal_context_net import ConvDiscriminator, PoseGenerator  # noqa: E402
from train_eval.train_gan import train_iter_gan  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/timeline.csv'
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
b = {k: v.to(dev) for k, v in bench.synth_batch(128, 1).items()}


def step():
    return train_iter_gan(args, 11, b['in_text'], b['in_audio'], b['target'], b['vid'], G, D, g_opt, d_opt)


for i in range(8):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        step()
    torch.cuda.synchronize()
prof.export_chrome_trace(out.replace('.csv', '.json'))
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start if evs else 0
os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
with open(out, 'w') as f:
    f.write('start_us,dur_us,stream,name\n')
    for e in evs:
        stream = getattr(e, 'stream', None)
        f.write('%.2f,%.2f,%s,"%s"\n' % (e.time_range.start - t0, e.time_range.end - e.time_range.start, stream, e.name[:90]))
print('events', len(evs), 'span_us', (evs[-1].time_range.end - t0) if evs else 0)
