import os, sys, time
ROOT='/root/repo'
sys.path.insert(0, ROOT)
import bench, torch
from model import vocab
from model.multimodal_context_net import ConvDiscriminator, PoseGenerator
from train_eval.staging import DevicePrefetcher
from train_eval.train_gan import train_iter_gan
dev = torch.device('cuda:0'); torch.manual_seed(0)
args = bench.make_args_ns()
spk = vocab.Vocab('vid', insert_default_tokens=False)
while spk.n_words < bench.N_SPEAKERS: spk.index_word('s%d' % spk.n_words)
G = PoseGenerator(args, bench.POSE_DIM, bench.N_WORDS, 300, None, z_obj=spk).to(dev).train()
D = ConvDiscriminator(bench.POSE_DIM).to(dev).train()
g_opt = torch.optim.Adam(G.parameters(), lr=5e-4, betas=(0.5, 0.999)); d_opt = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.999))
pinned = [{k: v.pin_memory() for k, v in bench.synth_batch(128, i).items()} for i in range(8)]
res = [{k: v.to(dev) for k, v in p.items()} for p in pinned]
def run(n, ctas):
    it = iter(DevicePrefetcher((pinned[i % 8] for i in range(n)), dev, copy_ctas=ctas)) if ctas else None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        b = next(it) if it else res[i % 8]
        train_iter_gan(args, 11, b['in_text'], b['in_audio'], b['target'], b['vid'], G, D, g_opt, d_opt)
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
run(8, 0)
for ctas in (0, 64, 16, 4, 0, 8, 32):
    print('copy_ctas %3d: %.3f ms/step' % (ctas, run(30, ctas)))
