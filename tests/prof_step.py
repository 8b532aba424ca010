"""Profiling driver (not a test): a few fast-mode train_iter_gan steps at the benchmark shape, for ncu captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (sets up sys.path for the package)
import torch  # noqa: E402

from model import vocab  # noqa: E402
from model.multimodal_context_net import ConvDiscriminator, PoseGenerator  # noqa: E402
from train_eval.train_gan import train_iter_gan  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
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
for i in range(steps):
    ret = train_iter_gan(args, 11, b['in_text'], b['in_audio'], b['target'], b['vid'], G, D, g_opt, d_opt)
torch.cuda.synchronize()
print(ret)
