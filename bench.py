#!/usr/bin/env python
"""bench.py - G+D adversarial training step of the trimodal gesture model (BASELINE.json configs[1]) on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference algorithm on the box's host cores)

One "step" = one train_iter_gan call (epoch 11 > loss_warmup: 3 generator forwards, 1 generator backward, 3
discriminator forward/backwards, two Adam updates) on a batch of 128 synthetic TED-shaped clips per GPU.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

N_WORDS, N_SPEAKERS, AUDIO_LEN, T, POSE_DIM = 20000, 1371, 36267, 34, 27
FLOP_PER_SAMPLE_STEP = 2.735e9          # 5 x 521.7 MFLOP (3 G fwd + G bwd@2x) + 9 x 14.05 MFLOP
FLOP_PER_CLIP_FWD = 521.7e6


def make_args_ns():
    return argparse.Namespace(n_pre_poses=4, n_poses=T, input_context='both', hidden_size=300, n_layers=4, dropout_prob=0.3,
                              freeze_wordembed=False, z_type='speaker', loss_warmup=10, loss_gan_weight=5.0,
                              loss_regression_weight=500.0, loss_kld_weight=0.1, loss_reg_weight=0.05, wordembed_dim=300)


def synth_batch(batch, seed):
    """Synthetic TED-shaped clips: audio 0.1*N(0,1) clipped, mostly-PAD word ids with 5-9 word frames,
    random-walk direction vectors, speaker ids."""
    rng = np.random.Generator(np.random.PCG64(seed))
    audio = np.clip(0.1 * rng.standard_normal((batch, AUDIO_LEN)), -1, 1).astype(np.float32)
    text = np.zeros((batch, T), dtype=np.int64)
    for b in range(batch):
        n = int(rng.integers(5, 10))
        text[b, rng.choice(T, size=n, replace=False)] = rng.integers(4, N_WORDS, size=n)
    walk = np.cumsum(0.02 * rng.standard_normal((batch, T, POSE_DIM)), axis=1)
    target = (walk + 0.1 * rng.standard_normal((batch, 1, POSE_DIM))).astype(np.float32)
    vid = rng.integers(1, N_SPEAKERS, size=batch).astype(np.int64)
    return dict(in_text=torch.from_numpy(text), in_audio=torch.from_numpy(audio), target=torch.from_numpy(target),
                vid=torch.from_numpy(vid))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sust=p['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU restatement of the reference algorithm, pinned to the reference's own
# modules by tests/test_oracle_golden.py) timed on the host cores.  The only place bench.py touches oracle/.
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, batch):
    """The oracle's G+D iteration on ALL host cores, always on the FULL per-GPU batch (never a self-chosen sample: round 1's probe shrank
    the batch on slow boxes and made the ratio incomparable).  ~1 s / step on 16 cores at batch 128."""
    from oracle import synth
    from oracle import trimodal_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.HotPathConfig(n_words=N_WORDS, n_speakers=N_SPEAKERS)
    gsd, dsd = synth.generator_state_dict(cfg), synth.discriminator_state_dict(cfg)
    g_opt, d_opt = synth.zeros_like_opt(gsd), synth.zeros_like_opt(dsd)

    def one(it):
        inp = synth_batch(batch, 100 + it)
        noise = synth.make_noise(cfg, batch, seed=it, dropout=True)
        t0 = time.perf_counter()
        O.train_iter_gan_oracle(cfg, 11, gsd, dsd, g_opt, d_opt, 1, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], noise)
        return time.perf_counter() - t0
    for i in range(warmup):
        one(i + 1)
    ts = [one(50 + i) for i in range(steps)]
    return dict(batch=batch, ms=1e3 * float(np.mean(ts)), cores=torch.get_num_threads())


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_reference_run(a.steps, a.warmup, a.batch)
    val = r['batch'] / (r['ms'] / 1e3)
    sample = 'G+D step (epoch 11, all dropout masks) on the full %d-clip batch, %d timed steps after %d warm-up' % (r['batch'], a.steps, a.warmup)
    line = {'impl': 'reference', 'metric': 'G+D train samples/s', 'value': val, 'unit': 'samples/s', 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': r['ms'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(a, 1),
            'cpu_baseline': {'value': val, 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def workload_config(a, world):
    return {'workload': 'multimodal_context G+D adversarial training step (train_iter_gan, epoch>loss_warmup), batch %d per GPU, '
                        '34 frames x 27-d poses, 4 seed poses, 36267 audio samples, 34-word ids, n_words=20000, 1370 speakers' % a.batch,
            'global_batch': a.batch * world, 'per_gpu_batch': a.batch, 'parallelism': 'dp%d' % world,
            'mode': os.environ.get('TGB200_MODE', 'tf32') + ' (tf32: tcgen05 tensor-core GEMM/GRU kernels, fp32 accumulate; fp32: CUDA-core FFMA kernels)',
            'l2': 'per-step working set (~1.5 GB of activations) exceeds the 126 MB L2; inputs rotate over 8 distinct batches'}


def fgd_workload(dev, world, rank):
    """BASELINE.json configs[4]: EmbeddingNet('pose') features of 10k real + 10k generated synthetic clips -> fp64 moments on the device ->
    Frechet distance.  The random-init net's BatchNorm running statistics are CALIBRATED first (20 train-mode batches through the
    stock-torch restatement, SURVEY 8d: without it the features barely vary and "FGD within 1 %" is meaningless), and the value is compared
    with the oracle's FGD (oracle encoder in fp64 on the same device, NumPy covariance + SciPy sqrtm).  With world > 1 every rank
    evaluates its own shard of the 10k pairs and the sufficient statistics are all-reduced (get_scores(reduce=True))."""
    from model.embedding_net import EmbeddingNet
    from model.embedding_space_evaluator import EmbeddingSpaceEvaluator
    from oracle import embed_train_oracle as EO
    from oracle import synth
    from oracle import trimodal_oracle as O
    cfg = O.HotPathConfig(n_words=N_WORDS, n_speakers=N_SPEAKERS)
    sd = {k: v.to(dev) for k, v in synth.embedding_net_state_dict(cfg).items()}
    g = torch.Generator(device='cpu').manual_seed(5)
    real = (0.5 * torch.randn(10000, T, POSE_DIM, generator=g)).to(dev)
    fake = (torch.randn(10000, T, POSE_DIM, generator=g) + 0.3).to(dev)
    # calibration: torch BatchNorm semantics (momentum 0.1, unbiased running variance) over 20 batches of 512 real-like clips
    gc = torch.Generator(device='cpu').manual_seed(6)
    with torch.no_grad():
        for _ in range(20):
            x = (0.5 * torch.randn(512, T, POSE_DIM, generator=gc)).to(dev)
            stats = {}
            EO.embedding_net_pose(sd, x, True, stats)          # train-mode forward of encoder + decoder: new running statistics
            sd.update(stats)
    e_args = argparse.Namespace(hidden_size=300, n_layers=4, dropout_prob=0.3, freeze_wordembed=False)
    enet = EmbeddingNet(e_args, POSE_DIM, T, N_WORDS, 300, None, 'pose')
    enet.load_state_dict({k: v.cpu() for k, v in sd.items()}, strict=True)
    enet = enet.to(dev)
    ev = EmbeddingSpaceEvaluator.from_net(enet, 4, dev)
    lo, hi = (10000 * rank) // world, (10000 * (rank + 1)) // world

    def fgd():
        ev.reset()
        for o in range(lo, hi, 500):
            e = min(hi, o + 500)
            ev.push_samples(None, None, fake[o:e], real[o:e])
        return ev.get_scores(reduce=world > 1)
    fgd()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    score = fgd()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {'fgd_10k_pairs_seconds': dt, 'fgd_clips_per_s': 20000 / dt, 'fgd_value': score[0], 'fgd_feat_dist': score[1], 'fgd_n_gpus': world}
    if rank == 0:
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        with torch.no_grad():
            rf = torch.cat([O.pose_encoder_conv(sd64, real[o:o + 2000].double()) for o in range(0, 10000, 2000)]).cpu().numpy()
            ff = torch.cat([O.pose_encoder_conv(sd64, fake[o:o + 2000].double()) for o in range(0, 10000, 2000)]).cpu().numpy()
        ref, ref_fd = O.fgd_scores(ff, rf)
        out['fgd_oracle'] = ref
        out['fgd_rel_err_vs_oracle'] = abs(score[0] - ref) / abs(ref)
        out['fgd_feat_dist_rel_err_vs_oracle'] = abs(score[1] - ref_fd) / abs(ref_fd)
        out['fgd_note'] = 'BatchNorm-calibrated random-init EmbeddingNet; oracle = fp64 encoder + np.cov + scipy sqrtm over all 10k pairs'
    return out


def aux_workloads(dev, timed, joint=False):
    """Side measurements of the other BASELINE.json configs on one GPU (reported next to the headline, not part of it):
    configs[3] seq2seq training step at batch 128; the FGD auto-encoder's training step."""
    out = {}
    from model.seq2seq_net import Seq2SeqNet
    from train_eval.train_seq2seq import train_iter_seq2seq
    s_args = argparse.Namespace(hidden_size=200, n_layers=2, dropout_prob=0.1, n_pre_poses=4, GAN_noise_size=0, loss_regression_weight=250.0,
                                loss_kld_weight=0.1, loss_reg_weight=25.0)                     # config/seq2seq.yml:17-19,26-28
    net = Seq2SeqNet(s_args, POSE_DIM, T, N_WORDS, 300, None).to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    rng = np.random.Generator(np.random.PCG64(7))
    lengths = np.sort(rng.integers(4, 13, size=128))[::-1].copy(); lengths[0] = 12
    text_np = np.zeros((128, 12), dtype=np.int64)
    for b in range(128):                                            # [SOS=1, words.., EOS=2], PAD 0, sorted by decreasing length
        n = int(lengths[b]); text_np[b, 0] = 1; text_np[b, 1:n - 1] = rng.integers(4, N_WORDS, size=n - 2); text_np[b, n - 1] = 2
    inp = {'lengths': torch.from_numpy(lengths.astype(np.int64))}
    text, target = torch.from_numpy(text_np).to(dev), synth_batch(128, 77)['target'].to(dev)
    f = lambda i: train_iter_seq2seq(s_args, 0, text, inp['lengths'], target, net, opt)
    for i in range(5):
        f(i)
    ms, _, _, _ = timed(f, 20)
    out['seq2seq_train_samples_per_s'] = 128 * 20 / (ms / 1e3)
    out['seq2seq_config'] = ('config/seq2seq.yml: hidden 200, 2 layers, batch 128, text length 4..12, 34 frames; %s mode (projections the TMA can '
                             'describe on tcgen05 TF32 tiles in tf32 mode, FFMA in fp32 mode), CUDA-graph replay' % os.environ.get('TGB200_MODE', 'tf32'))
    # training step of the FGD auto-encoder (train_feature_extractor.train_iter), batch 128, CUDA-graph replay
    try:
        from model.embedding_net import EmbeddingNet
        import train_feature_extractor as tfx
        e_args = argparse.Namespace(hidden_size=300, n_layers=4, dropout_prob=0.3, freeze_wordembed=False)
        anet = EmbeddingNet(e_args, POSE_DIM, T, None, None, None, 'pose').to(dev).train()
        aopt = torch.optim.Adam(anet.parameters(), lr=5e-4, betas=(0.5, 0.999))
        tg = [synth_batch(128, 90 + i)['target'].to(dev) for i in range(4)]
        fa = lambda i: tfx.train_iter(None, 0, tg[i % 4], anet, aopt)
        for i in range(5):
            fa(i)
        ms, _, _, _ = timed(fa, 50)
        out['autoencoder_train_samples_per_s'] = 128 * 50 / (ms / 1e3)
        out['autoencoder_train_ms_per_step'] = ms / 50
    except Exception as exc:                                          # a side measurement must never take the headline line down
        out['autoencoder_train_error'] = '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0] if str(exc) else '')
    # training step of the Speech2Gesture baseline (train_iter_speech2gesture), batch 128, eager launches
    try:
        from model.speech2gesture import Discriminator as S2GD, Generator as S2GG
        from train_eval.train_speech2gesture import train_iter_speech2gesture
        sg, sd_ = S2GG(T, POSE_DIM, 4).to(dev).train(), S2GD(POSE_DIM).to(dev).train()
        sgo = torch.optim.Adam(sg.parameters(), lr=1e-3, betas=(0.5, 0.999)); sdo = torch.optim.Adam(sd_.parameters(), lr=2e-4, betas=(0.5, 0.999))
        sa = argparse.Namespace(n_pre_poses=4, loss_regression_weight=100.0, loss_gan_weight=10.0)
        gsp = torch.Generator().manual_seed(5)
        spec = (torch.randn(128, 128, 70, generator=gsp) * 20.0 - 40.0).to(dev)
        tg2 = [synth_batch(128, 120 + i)['target'].to(dev) for i in range(2)]
        fs = lambda i: train_iter_speech2gesture(sa, spec, tg2[i % 2], sg, sd_, sgo, sdo, None)
        for i in range(3):
            fs(i)
        ms, _, _, _ = timed(fs, 10)
        out['speech2gesture_train_samples_per_s'] = 128 * 10 / (ms / 1e3)
        out['speech2gesture_train_ms_per_step'] = ms / 10
        del sg, sd_, sgo, sdo
        torch.cuda.empty_cache()
    except Exception as exc:
        out['speech2gesture_train_error'] = '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0] if str(exc) else '')
    if not joint:
        return out
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'bench_joint.py')], capture_output=True, text=True, timeout=180, cwd=ROOT)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
        if r.returncode == 0 and line:
            j = json.loads(line[-1])
            out['joint_embed_train_samples_per_s'] = j['tf32']['samples_per_s']
            out['joint_embed_train'] = j
        else:
            out['joint_embed_train_error'] = (r.stderr.strip().splitlines() or ['rc=%d' % r.returncode])[-1][:300]
    except Exception as exc:
        out['joint_embed_train_error'] = '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0] if str(exc) else '')
    return out


def stock_torch_baseline(dev, batch, timed, steps):
    """The same G+D iteration through STOCK PyTorch on this GPU (cuDNN convolutions, cuDNN RNN, cuBLAS; oracle/stock_torch.py, pinned to
    the oracle by tests/test_stock_torch_pinned.py): the 'existing Blackwell kernels' bar of SURVEY 2.3 / 8d.  fp32 (TF32 off), TF32
    allowed, and bf16 autocast; eager, like the reference runs it."""
    from oracle import stock_torch as ST
    from oracle import synth
    from oracle import trimodal_oracle as O
    cfg = O.HotPathConfig(n_words=N_WORDS, n_speakers=N_SPEAKERS)
    gsd, dsd = synth.generator_state_dict(cfg), synth.discriminator_state_dict(cfg)
    data = [{k: v.to(dev) for k, v in synth_batch(batch, 500 + i).items()} for i in range(4)]
    out = {'what': 'oracle/stock_torch.py: nn.Conv1d / nn.GRU / weight_norm modules + autograd + torch.optim.Adam, eager, batch %d' % batch}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for tag, tf32, amp in (('fp32', False, False), ('tf32', True, False), ('bf16_autocast', True, True)):
        try:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            G, D, g_opt, d_opt = ST.build(cfg, gsd, dsd, dev)

            def f(i):
                b = data[i % 4]
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
                    return ST.train_iter_gan_stock(cfg, 11, b['in_text'], b['in_audio'], b['target'], b['vid'], G, D, g_opt, d_opt)
            for i in range(4):
                ret = f(i)
            ms, _, _, _ = timed(f, steps)
            out[tag] = {'samples_per_s': batch * steps / (ms / 1e3), 'ms_per_step': ms / steps, 'finite': bool(all(np.isfinite(v) for v in ret.values()))}
            del G, D, g_opt, d_opt
        except Exception as exc:
            out[tag] = {'error': '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0] if str(exc) else '')}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()
    return out


def percentiles(xs):
    xs = np.asarray(xs, dtype=np.float64)
    return {'p10': float(np.percentile(xs, 10)), 'p50': float(np.percentile(xs, 50)), 'p90': float(np.percentile(xs, 90)), 'max': float(xs.max())}


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=128, help='clips per GPU')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-profile', action='store_true')
    ap.add_argument('--aux-joint', action='store_true', help='also time the joint-embedding training step (child process, tests/bench_joint.py)')
    ap.add_argument('--no-aux', action='store_true', help='skip the seq2seq / FGD / auto-encoder side measurements (BASELINE.json configs[3], configs[4])')
    ap.add_argument('--no-modes', action='store_true', help='skip the second arithmetic mode (strict fp32) of BASELINE.json configs[1]')
    ap.add_argument('--no-stock', action='store_true', help='skip the stock-PyTorch-on-this-GPU baseline leg')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling point (global batch 1024, BASELINE.json configs[2])')
    ap.add_argument('--e2e-steps', type=int, default=100)
    a = ap.parse_args()
    if a.impl == 'reference':
        return run_reference(a)
    a.warmup = max(a.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs CUDA devices; there is no CPU fallback for the product path'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')            # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group('nccl', device_id=dev)

    from model import vocab
    from model.multimodal_context_net import ConvDiscriminator, PoseGenerator
    from tgb200 import config as tg_config
    from tgb200 import ops
    from tgb200.profiler import KernelTimer
    from train_eval.staging import DevicePrefetcher
    from train_eval.train_gan import train_iter_gan

    args = make_args_ns()
    spk = vocab.Vocab('vid', insert_default_tokens=False)
    while spk.n_words < N_SPEAKERS:
        spk.index_word('s%d' % spk.n_words)

    def build_models():
        torch.manual_seed(rank)                            # ranks start from DIFFERENT weights: the first data-parallel step broadcasts rank 0's
        G = PoseGenerator(args, POSE_DIM, N_WORDS, 300, None, z_obj=spk).to(dev).train()
        D = ConvDiscriminator(POSE_DIM).to(dev).train()
        g_opt = torch.optim.Adam(G.parameters(), lr=5e-4, betas=(0.5, 0.999))
        d_opt = torch.optim.Adam(D.parameters(), lr=5e-4 * 0.2, betas=(0.5, 0.999))
        return G, D, g_opt, d_opt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, per_step=False):
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1 if per_step else 2)]
        l0 = ops.launches()
        t0 = time.time()
        evs[0].record()
        for i in range(steps):
            fn(i)
            if per_step:
                evs[i + 1].record()
        if not per_step:
            evs[1].record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([evs[0].elapsed_time(evs[-1])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if per_step:
            return ms.item(), [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        return ms.item(), ops.launches() - l0, t0, t1

    n_pool = 8

    def make_pool(batch):
        host = [synth_batch(batch, 1000 * rank + i) for i in range(n_pool)]
        return host, [{k: v.to(dev) for k, v in h.items()} for h in host]

    def resident_run(G, D, g_opt, d_opt, pool, steps, warm):
        def step(i):
            b = pool[i % n_pool]
            return train_iter_gan(args, 11, b['in_text'], b['in_audio'], b['target'], b['vid'], G, D, g_opt, d_opt)
        l0 = ops.launches()
        ret = step(0)                                      # first (eager) iteration: counts the kernel launches of one step
        lps = ops.launches() - l0
        for i in range(1, max(warm, 4)):                   # >= 2 eager iterations, then the CUDA graph is captured and replayed
            ret = step(i)
        assert all(np.isfinite(v) for v in ret.values()), ret
        return step, lps, ret

    # ---- headline: BASELINE.json configs[1] in the mode of TGB200_MODE (default tf32), inputs resident in HBM
    mode0 = tg_config.mode()
    G, D, g_opt, d_opt = build_models()
    host, resident = make_pool(a.batch)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    step_resident, launches_per_step, ret = resident_run(G, D, g_opt, d_opt, resident, a.steps, a.warmup)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    ms, launches, t0, t1 = timed(step_resident, a.steps)
    clock_info = clocks.stop(t0, t1) if rank == 0 else None
    value = world * a.batch * a.steps / (ms / 1e3)

    # ---- end to end: the timer starts BEFORE the prefetcher exists; every step's inputs come from pinned host memory inside the timed
    # region (train_eval.staging.DevicePrefetcher copies batch i+1 on its own stream while step i runs; the reference does a blocking
    # .to(device) per batch, train.py:171-176) and the step's result dict is read back (D2H).  >= 100 steps, per-step percentiles.
    pinned = [{k: v.pin_memory() for k, v in h.items()} for h in host]
    e2e_steps = max(a.e2e_steps, a.steps)

    def e2e_run(steps):
        state = {}

        def f(i):
            h0 = time.perf_counter()
            if i == 0:
                state['it'] = iter(DevicePrefetcher((pinned[j % n_pool] for j in range(steps)), dev))
            b = next(state['it'])
            h1 = time.perf_counter()
            r = train_iter_gan(args, 11, b['in_text'], b['in_audio'], b['target'], b['vid'], G, D, g_opt, d_opt)   # returns python floats (D2H)
            if i == 0:
                state['first'] = {'prefetcher_ctor_and_first_batch_host_ms': 1e3 * (h1 - h0), 'first_step_host_ms': 1e3 * (time.perf_counter() - h1)}
            return r
        out = timed(f, steps, per_step=True)
        return out[0], out[1], state.get('first')
    e2e_run(3)
    ms_e2e, per_step_e2e, e2e_first = e2e_run(e2e_steps)
    e2e_value = world * a.batch * e2e_steps / (ms_e2e / 1e3)
    # diagnostics for a slow box: one batch's host->device copy alone, through the same prefetcher path
    pf = DevicePrefetcher([pinned[0]] * 6, dev)
    copy_ms = []
    for _ in pf:
        torch.cuda.synchronize()
    ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ce0.record(pf.stream)
    with torch.cuda.stream(pf.stream):
        for k, v in pinned[1].items():
            if v.numel() * v.element_size() >= (1 << 20):
                ops.copy_bytes(resident[1][k], v, pf.copy_ctas)
    ce1.record(pf.stream)
    torch.cuda.synchronize()
    copy_ms = ce0.elapsed_time(ce1)

    # ---- PoseGenerator inference (the metric's "clips/s"): eval forward, batch 512 / 128 / 1, device-resident inputs
    G.eval()
    infer = {}
    with torch.no_grad():
        for bs in (512, a.batch, 1):
            b = {k: torch.cat([r[k] for r in resident[:(bs + a.batch - 1) // a.batch]])[:bs].contiguous() for k in resident[0]}
            pre = torch.zeros(bs, T, POSE_DIM + 1, device=dev)
            pre[:, :4, :-1] = b['target'][:, :4]; pre[:, :4, -1] = 1
            f = lambda i: G(pre, b['in_text'], b['in_audio'], b['vid'])
            for i in range(3):
                f(i)
            n_it = 20
            ims, _, _, _ = timed(f, n_it)
            infer['b%d' % bs] = world * bs * n_it / (ims / 1e3)
    G.train()

    # ---- per-kernel pass (dominant kernel, family table) on the headline models
    kt = None
    if not a.no_kernel_profile:
        # per-launch CUDA-event timing of two more steps.  EVERY rank runs them (the steps contain the gradient all-reduces: a rank-0-only
        # pass would wait for its peers forever); only rank 0 reports.
        old_graphs = tg_config.set_graphs(False)      # per-launch events need eager launches
        old_overlap = tg_config.set_overlap(False)    # one stream: an event pair must not include waiting for SMs held by another stream
        with KernelTimer() as kt:
            for i in range(2):
                # park the GPU behind ~25 ms of spin so the host (slower than the GPU in eager mode) queues the whole step
                # ahead of it: every event pair then brackets exactly one kernel's device time, not the host's launch gap
                torch.cuda._sleep(int(0.025 * 1.9e9))
                step_resident(i)
        tg_config.set_graphs(old_graphs)
        tg_config.set_overlap(old_overlap)
        barrier()

    # ---- BASELINE.json configs[1] "fp32 and bf16": the other arithmetic mode on fresh models (bf16 -> this repo's tf32 mode)
    modes = {mode0: {'samples_per_s': value, 'ms_per_step': ms / a.steps}}
    if not a.no_modes and world == 1:
        other = 'fp32' if mode0 == 'tf32' else 'tf32'
        old_mode = tg_config.set_mode(other)
        try:
            G2, D2, g2, d2 = build_models()
            step2, _, _ = resident_run(G2, D2, g2, d2, resident, a.steps, a.warmup)
            ms2, _, _, _ = timed(step2, a.steps)
            modes[other] = {'samples_per_s': a.batch * a.steps / (ms2 / 1e3), 'ms_per_step': ms2 / a.steps}
            del G2, D2, g2, d2
        except Exception as exc:
            modes[other] = {'error': '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0] if str(exc) else '')}
        tg_config.set_mode(old_mode)

    # ---- strong scaling (BASELINE.json configs[2]): global batch 1024 sharded over the ranks (1024 / N clips per GPU)
    strong = None
    if not a.no_strong:
        sb = 1024 // world
        if sb == a.batch:
            strong = {'global_batch': 1024, 'per_gpu_batch': sb, 'value': value, 'ms_per_step': ms / a.steps, 'note': 'same shape as the weak-scaling headline'}
        else:
            try:
                _, res_s = make_pool(sb)
                G3, D3, g3, d3 = build_models()
                step3, _, _ = resident_run(G3, D3, g3, d3, res_s, 8, 4)
                ms3, _, _, _ = timed(step3, 8)
                strong = {'global_batch': 1024, 'per_gpu_batch': sb, 'value': world * sb * 8 / (ms3 / 1e3), 'ms_per_step': ms3 / 8, 'steps': 8}
                del G3, D3, g3, d3, res_s
                torch.cuda.empty_cache()
            except Exception as exc:
                strong = {'global_batch': 1024, 'per_gpu_batch': sb, 'error': '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0] if str(exc) else '')}

    aux = {}
    if not a.no_aux:
        try:
            aux.update(fgd_workload(dev, world, rank))
        except Exception as exc:
            aux['fgd_error'] = '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0] if str(exc) else '')
        if world == 1:
            aux.update(aux_workloads(dev, timed, joint=a.aux_joint))
    stock = None
    if world == 1 and not a.no_stock:
        stock = stock_torch_baseline(dev, a.batch, timed, 10)

    line = None
    if rank == 0:
        peaks = load_peaks()
        roof = None
        if kt is not None:
            agg = kt.summary()
            try:
                os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
                with open(os.path.join(ROOT, 'gpurun_out', 'kernels_by_shape.txt'), 'w') as f:
                    for (kn, tag), x in sorted(kt.by_shape().items(), key=lambda kv: -kv[1]['ms']):
                        f.write('%-26s %-34s calls/step %5.1f  ms/step %7.4f  TFLOP/s %7.1f\n' % (
                            kn, tag, x['calls'] / 2, x['ms'] / 2, (x['flops'] / (x['ms'] / 1e3) / 1e12) if x['ms'] > 0 else 0.0))
            except OSError:
                pass
            total_ms = sum(v['ms'] for v in agg.values())
            # dominant kernel = the (entry point, shape) with the most device time per step; the family table below shows the entry points
            (name, tag), v = max(kt.by_shape().items(), key=lambda kv: kv[1]['ms'])
            tflops = v['flops'] / (v['ms'] / 1e3) / 1e12 if v['ms'] > 0 else 0.0
            traffic = None
            try:                                            # dram__bytes_read+write per launch from the committed ncu --set full capture
                for fn in ('r02_ncu_traffic.json', 'r01_ncu_traffic.json'):
                    pth = os.path.join(ROOT, 'profiles', fn)
                    if os.path.exists(pth):
                        traffic = json.load(open(pth)).get('%s|%s' % (name, tag), {}).get('dram_bytes_per_launch')
                        if traffic is not None:
                            break
            except (OSError, ValueError):
                pass
            kinds = {'tg_gemm_tf32': 'tcgen05 kind::tf32 GEMM family (all shapes of a step, fp32 accumulate in TMEM)',
                     'tg_wgrad_tf32': 'tcgen05 kind::tf32 weight-gradient GEMM family',
                     'tg_gru_layer_fwd_tf32': '8-CTA-cluster tensor-core GRU forward (latency-bound recurrence, 34 sequential steps per launch)',
                     'tg_gru_layer_bwd_tf32': '8-CTA-cluster tensor-core GRU backward (latency-bound recurrence)'}
            families = {}
            for k, x in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])[:14]:
                fam = {'ms_per_step': round(x['ms'] / 2, 4), 'launches_per_step': x['calls'] / 2, 'share_of_kernel_time': round(x['ms'] / total_ms, 4)}
                if x['flops'] > 0 and x['ms'] > 0:
                    fam['tflops'] = round(x['flops'] / (x['ms'] / 1e3) / 1e12, 2)
                    fam['frac_of_bf16_peak'] = round(fam['tflops'] / peaks['tf_sust'], 4)
                elif x['bytes'] > 0 and x['ms'] > 0:
                    fam['gbs'] = round(x['bytes'] / (x['ms'] / 1e3) / 1e9, 1)
                    fam['frac_of_hbm_peak'] = round(fam['gbs'] / peaks['hbm'], 4)
                families[k] = fam
            roof = {'bound': 'tensor', 'kernel': name, 'shape': tag, 'achieved': tflops, 'peak': peaks['tf_sust'], 'unit': 'TFLOP/s',
                    'frac': tflops / peaks['tf_sust'], 'traffic': traffic, 'peak_source': peaks['src'] + ' bf16 dense sustained',
                    'share_of_step': v['ms'] / total_ms, 'launches_per_step': v['calls'] / 2,
                    'avg_launch_us': 1e3 * v['ms'] / v['calls'], 'flops_per_launch': v['flops'] / v['calls'],
                    'note': '%s; algorithmic FLOPs of its launches / their summed CUDA-event durations, measured against the bf16 tensor-pipe '
                            'peak (TF32 peak is half of it); whole-step fraction = %.4f'
                            % (kinds.get(name, 'fp32 CUDA-core kernel'), value / world * FLOP_PER_SAMPLE_STEP / (peaks['tf_sust'] * 1e12)),
                    'families': families}
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            r = cpu_reference_run(2, 1, a.batch)
            cpu = {'value': r['batch'] / (r['ms'] / 1e3), 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port',
                   'sample': 'oracle G+D step on the full %d-clip batch, 2 timed steps after 1 warm-up' % r['batch']}
        line = {'metric': 'G+D train samples/s', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
                'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'tf32' if mode0 == 'tf32' else 'f32', 'data': 'synthetic',
                'config': workload_config(a, world),
                'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 64,
                        'ms_per_step': ms_e2e / e2e_steps, 'steps': e2e_steps, 'per_step_ms': percentiles(per_step_e2e),
                        'one_batch_h2d_ms': copy_ms, 'first_step': e2e_first, 'timer': 'starts before the prefetcher is constructed'},
                'gpu_launches': launches_per_step * a.steps, 'launches_per_step': launches_per_step,
                'cuda_graph': tg_config.graphs(), 'clocks': clock_info, 'roofline': roof, 'cpu_baseline': cpu,
                'step_roofline_frac': value / world * FLOP_PER_SAMPLE_STEP / (peaks['tf_sust'] * 1e12),
                'modes': modes, 'strong_scaling': strong, 'gpu_stock_baseline': stock,
                'infer_clips_per_s': infer, 'aux': aux, 'last_losses': ret}
        print(json.dumps(line))
    if world > 1:
        # A captured CUDA graph that contains NCCL kernels keeps the communicator busy: destroy_process_group() behind it hung the process
        # until the job's timeout (round-2 call L).  Order: drain the device, barrier, flush the one JSON line, leave without the teardown.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
