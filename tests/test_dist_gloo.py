"""CPU, world_size 2, gloo: the data-parallel host logic of train_iter_gan - bucketed all-reduce of the flat gradient
arena (one process per GPU in production, NCCL over NVLink; here the same code path over gloo) and the sharded FGD
statistics reduction.  No kernel is launched."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gpu_util import build_ours
        from oracle.make_golden import golden_cfg
        from tgb200.engine import gru_arena_order
        from tgb200.arena import ParamArena
        from train_eval.train_gan import _allreduce_grads, _dist_world
        assert _dist_world() == world
        cfg = golden_cfg()
        args, G, D, _, _ = build_ours(cfg, None)
        arena = ParamArena(D, gru_arena_order([n for n, _ in D.named_parameters()])).ensure(torch.device('cpu'))
        g = torch.Generator().manual_seed(100 + rank)
        arena.grad.copy_(torch.randn(arena.numel, generator=g))
        mine = arena.grad.clone()
        _allreduce_grads(arena, bucket_floats=10007)          # several ragged buckets
        others = [torch.randn(arena.numel, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        expect = sum(others)
        ok = torch.allclose(arena.grad, expect, atol=1e-5) and torch.allclose(others[rank], mine)
        # .grad views of the parameters see the reduced values (they alias the flat buffer)
        p = dict(D.named_parameters())['gru.weight_hh_l2']
        o = arena.offsets['gru.weight_hh_l2']
        ok = ok and torch.equal(p.grad.reshape(-1), arena.grad[o:o + p.numel()])
        # FGD: per-rank sufficient statistics -> all-reduce -> identical scores on every rank
        from model.embedding_space_evaluator import EmbeddingSpaceEvaluator
        ev = EmbeddingSpaceEvaluator.__new__(EmbeddingSpaceEvaluator)
        ev.device = torch.device('cpu'); ev.reset()
        feats_r = torch.randn(64, 32, generator=torch.Generator().manual_seed(7 + rank)).double()
        feats_g = (torch.randn(64, 32, generator=torch.Generator().manual_seed(17 + rank)) * 1.3 + 0.2).double()
        for acc, f in ((ev.acc_real, feats_r), (ev.acc_gen, feats_g)):
            acc[0] = f.shape[0]; acc[1:33] = f.sum(0); acc[33:] = (f.t() @ f).reshape(-1)
        ev.acc_misc[0] = (feats_r - feats_g).abs().sum()
        fgd, fdist = ev.get_scores(reduce=True)
        q.put((rank, bool(ok), float(fgd), float(fdist)))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_and_fgd_reduction_world2():
    import numpy as np
    from oracle import trimodal_oracle as O
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert abs(res[0][2] - res[1][2]) < 1e-9 and abs(res[0][3] - res[1][3]) < 1e-9
    # the reduced FGD equals the oracle's FGD over the union of both ranks' features
    fr = np.vstack([torch.randn(64, 32, generator=torch.Generator().manual_seed(7 + r)).double().numpy() for r in range(2)])
    fg = np.vstack([(torch.randn(64, 32, generator=torch.Generator().manual_seed(17 + r)) * 1.3 + 0.2).double().numpy() for r in range(2)])
    fgd, fdist = O.fgd_scores(fg, fr)
    assert abs(res[0][2] - fgd) <= 1e-6 * abs(fgd) and abs(res[0][3] - fdist) <= 1e-9 * abs(fdist)


# ----------------------------------------------------------------------------------------------------------------------
# Full data-parallel iteration, world 2: every rank runs train_iter_gan on ITS shard (the launch plan executes on the NumPy C-ABI
# emulator, tests/cabi_emulator.py), the flat gradient arenas are summed over gloo where production uses NCCL, Adam averages.
# ----------------------------------------------------------------------------------------------------------------------
def _dp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(4)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import cabi_emulator
        from gpu_util import build_ours, masks_to_ours
        from oracle import synth
        from oracle import trimodal_oracle as O
        from oracle.make_golden import golden_cfg
        from tgb200 import config
        from train_eval import train_gan as TG
        config.set_mode('fp32'); config.set_graphs(False)
        cfg = golden_cfg()
        Bl = 2
        full = synth.make_inputs(cfg, Bl * world, seed=9)
        inp = {k: v[rank * Bl:(rank + 1) * Bl].contiguous() for k, v in full.items()}
        res = {}
        with cabi_emulator.installed() as emu:
            for epoch in (0, 11):
                args, G, D, gsd, dsd = build_ours(cfg, None)
                G.train(); D.train()
                noise = synth.golden_noise(cfg, Bl, 20 + rank, True)          # every rank draws its own noise
                g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
                d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
                TG.inject_noise(TG.StepNoise(eps=list(noise.eps), perm=noise.perm, g_masks=[masks_to_ours(m, None) if m else {} for m in noise.g_masks],
                                             d_masks=[{}, {}, {}]))
                ret = TG.train_iter_gan(args, epoch, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)
                ok = all(torch.isfinite(torch.tensor(v)) for v in ret.values())
                # (1) replicas stay in lock-step: bit-identical parameters on every rank after the step
                for net in (G, D):
                    flat = net.engine().arena.flat.clone()
                    ref = flat.clone()
                    dist.broadcast(ref, src=0)
                    ok = ok and torch.equal(flat, ref)
                if epoch == 0:
                    # (2) warm-up epoch (no D step): the update equals Adam on the MEAN over ranks of the rank-local oracle gradients
                    from test_oracle_golden import ZERO_GRAD_KEYS
                    n3 = O.StepNoise(eps=list(noise.eps), perm=noise.perm, g_masks=list(noise.g_masks), d_masks=[None, None, None])
                    want = O.train_iter_gan_oracle(cfg, 0, gsd, dsd, synth.zeros_like_opt(gsd), synth.zeros_like_opt(dsd), 1,
                                                   inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], n3)
                    worst = 0.0
                    sd = G.state_dict()
                    for k, gr in want['g_grads'].items():
                        g = gr.detach().clone()
                        dist.all_reduce(g)
                        g /= world
                        p, _, _ = O.adam_step(gsd[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, cfg.learning_rate)
                        d = (sd[k] - p).abs()
                        ok = ok and d.max().item() <= 2.2 * cfg.learning_rate + 1e-6          # round-off gradients may flip a +-lr step
                        if k not in ZERO_GRAD_KEYS:                                            # analytically zero gradient: pure +-lr noise
                            worst = max(worst, d.median().item())
                    ok = ok and worst < 2e-6
                    res['worst_median'] = worst
                res['calls_%d' % epoch] = emu.calls.count('tg_adam_flat')
                res['ok_%d' % epoch] = bool(ok)
                emu.calls.clear()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_data_parallel_train_iter_gan_world2():
    for rank, r in _run_world2(_dp_worker, 30500):
        assert r['ok_0'] and r['ok_11'], (rank, r)
        # warm-up: generator Adam only (two launches: the recurrent range early, the rest at the end); afterwards D and G
        assert r['calls_0'] == 2 and r['calls_11'] == 3, r


# ----------------------------------------------------------------------------------------------------------------------
# The embedding-model trainers under data parallelism, world 2 (plans on the emulator, gradients over gloo)
# ----------------------------------------------------------------------------------------------------------------------
def _guarded(body, rank, world, port, q):
    try:
        body(rank, world, port, q)
    except BaseException as exc:          # report instead of leaving the peer blocked in a collective until the timeout
        import traceback
        q.put((rank, {'error': ''.join(traceback.format_exception(type(exc), exc, exc.__traceback__))[-1500:]}))
        os._exit(1)


def _run_world2(body, port_base, timeout=400):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = port_base + (os.getpid() % 1000)
    procs = [ctx.Process(target=_guarded, args=(body, r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=timeout))
            assert not (isinstance(res[-1][1], dict) and 'error' in res[-1][1]), res[-1][1]['error']
    finally:
        for p in procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
    return sorted(res, key=lambda t: t[0])


def _embed_dp_worker_body(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'gesture-generation-from-trimodal-context_b200'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(4)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import random
        import ae_checks
        import cabi_emulator
        import joint_checks
        import train_feature_extractor as tfx
        from oracle import embed_train_oracle as EO
        from oracle import synth
        from oracle import trimodal_oracle as O
        from tgb200 import config
        from train_eval.train_joint_embed import train_iter_embed
        config.set_mode('fp32'); config.set_graphs(False)
        res = {}
        with cabi_emulator.installed():
            # ---- auto-encoder: update == Adam on the mean over ranks of the rank-local oracle gradients; replicas identical
            cfg, net, opt = ae_checks.build(torch.device('cpu'))
            net.train()
            tgt = synth.make_inputs(cfg, 8, seed=33)['target'][rank * 4:(rank + 1) * 4].contiguous()
            ret = tfx.train_iter(None, 0, tgt, net, opt)
            sd0 = synth.embedding_net_state_dict(cfg)
            want = EO.train_iter_ae_oracle(sd0, synth.zeros_like_opt(sd0), 1, tgt, 5e-4, True)
            ok = abs(ret['loss'] - want['loss']) < 2e-5 * abs(want['loss'])          # the logged loss is rank-local
            sd = net.state_dict()
            worst = 0.0
            for k, gr in want['grads'].items():
                if k in EO.UNUSED_PARAMS:
                    continue
                g = gr.detach().clone()
                dist.all_reduce(g)
                g /= world
                p, _, _ = O.adam_step(sd0[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, 5e-4)
                d = (sd[k] - p).abs()
                ok = ok and d.max().item() <= 2.2 * 5e-4 + 1e-6
                if k not in EO.ZERO_GRAD_PARAMS:
                    worst = max(worst, d.median().item())
            ok = ok and worst < 2e-6
            flat = net.train_engine().arena.flat.clone(); ref = flat.clone()
            dist.broadcast(ref, src=0)
            res['ae'] = bool(ok and torch.equal(flat, ref))
            # ---- joint embedding, mode='random': ranks flip different coins, rank 0's choice is broadcast; replicas stay identical
            random.seed(100 + rank)
            cfg, args, jnet, jopt = joint_checks.build(torch.device('cpu'))
            jnet.train()
            same = True
            for step in range(3):
                data = synth.make_inputs(cfg, 4, seed=40 + step)
                data = {k: v[rank * 2:(rank + 1) * 2].contiguous() for k, v in data.items()}
                r = train_iter_embed(args, 0, data['in_text'], data['in_audio'], data['target'], jnet, jopt, mode='random')
                same = same and r['loss'] == r['loss']
                for a in jnet.joint_engine().arenas():
                    flat = a.flat.clone(); ref = flat.clone()
                    dist.broadcast(ref, src=0)
                    same = same and torch.equal(flat, ref)
            res['joint'] = bool(same)
            # ---- seq2seq: clip_grad_norm_ acts on the MEAN gradient: update == Adam(clip(mean over ranks of the rank-local raw gradients))
            import test_gpu_seq2seq as GS
            from oracle import seq2seq_oracle as S
            from train_eval.train_seq2seq import train_iter_seq2seq
            s_cfg = S.Seq2SeqConfig(n_words=200)
            s_args, s_net = GS._build(s_cfg, torch.device('cpu'))
            s_net.train()
            s_opt = torch.optim.Adam(s_net.parameters(), lr=s_cfg.learning_rate, betas=(0.9, 0.999))
            inp = synth.seq2seq_inputs(s_cfg, 4, seed=60 + rank, max_len=7)
            ssd = synth.seq2seq_state_dict(s_cfg)
            ref = S.train_iter_seq2seq_oracle(s_cfg, ssd, inp['in_text'], inp['lengths'], inp['target'], None, step=1)
            train_iter_seq2seq(s_args, 0, inp['in_text'], inp['lengths'], inp['target'], s_net, s_opt)
            keys = [k for k in ref['grads']]
            unclip = min(1.0, 5.0 / (float(ref['total_norm']) + 1e-6))
            mean = {}
            for k in keys:
                g = (ref['grads'][k] / unclip).clone()           # the oracle returns the rank's CLIPPED gradient: undo its own clipping
                dist.all_reduce(g)
                mean[k] = g / world
            norm = torch.sqrt(sum((g.double() ** 2).sum() for g in mean.values())).item()
            coef = min(1.0, 5.0 / (norm + 1e-6))
            ok = True
            sd2 = s_net.state_dict()
            worst = 0.0
            for k in keys:
                g = mean[k] * coef
                p, _, _ = O.adam_step(ssd[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, s_cfg.learning_rate, 0.9, 0.999)
                d = (sd2[k] - p).abs()
                ok = ok and d.max().item() <= 2.2 * s_cfg.learning_rate + 1e-6
                if float(ref['grads'][k].norm()) > 1e-6:
                    worst = max(worst, d.median().item())
            flat = s_net.engine().arena.flat.clone(); refp = flat.clone()
            dist.broadcast(refp, src=0)
            res['seq2seq'] = bool(ok and worst < 2e-6 and torch.equal(flat, refp))
            res['seq2seq_worst'] = worst
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_data_parallel_embedding_trainers_world2():
    for rank, r in _run_world2(_embed_dp_worker_body, 31500):
        assert r['ae'] and r['joint'] and r['seq2seq'], (rank, r)
