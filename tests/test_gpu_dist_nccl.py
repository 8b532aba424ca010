"""Data-parallel train_iter_gan on REAL hardware: 2 ranks, NCCL, the CUDA kernels (replaces nn.DataParallel, train.py:93-96).
Skipped with fewer than 2 GPUs (run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist_nccl.py -m gpu`).

Every rank runs the iteration on ITS shard with its own injected noise; afterwards
  (1) the replicas are bit-identical (flat parameter arenas of G and D compared across ranks),
  (2) the update equals Adam on the MEAN over ranks of the rank-local fp64-oracle gradients (warm-up epoch: generator only),
  (3) ranks that start from DIFFERENT weights end up identical (the first step broadcasts rank 0's parameters),
  (4) several CUDA-graph-replayed steps keep the replicas bit-identical (device-drawn, rank-distinct noise)."""
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
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    res = {}
    try:
        from gpu_util import build_ours, masks_to_ours, to_dev
        from oracle import synth
        from oracle import trimodal_oracle as O
        from oracle.make_golden import golden_cfg
        from test_oracle_golden import ZERO_GRAD_KEYS
        from tgb200 import config
        from train_eval import train_gan as TG
        cfg = golden_cfg()
        Bl = 4
        full = synth.make_inputs(cfg, Bl * world, seed=9)
        inp_cpu = {k: v[rank * Bl:(rank + 1) * Bl].contiguous() for k, v in full.items()}
        inp = to_dev(inp_cpu, dev)

        def same_across_ranks(net):
            flat = net.engine().arena.flat.clone()
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            return bool(torch.equal(flat, ref))

        for mode in ('fp32', 'tf32'):
            config.set_mode(mode); config.set_graphs(False)
            args, G, D, gsd, dsd = build_ours(cfg, dev)
            G.train(); D.train()
            noise = synth.golden_noise(cfg, Bl, 20 + rank, True)              # every rank draws its own noise
            g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
            d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
            TG.inject_noise(TG.StepNoise(eps=[e.to(dev) for e in noise.eps], perm=noise.perm.to(dev),
                                         g_masks=[masks_to_ours(m, dev) if m else {} for m in noise.g_masks], d_masks=[{}, {}, {}]))
            ret = TG.train_iter_gan(args, 0, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)
            ok = all(torch.isfinite(torch.tensor(v)) for v in ret.values()) and same_across_ranks(G) and same_across_ranks(D)
            # Adam on the mean of the rank-local oracle gradients (fp64 oracle on this rank's GPU)
            f64 = lambda sd: {k: (v.to(dev).double() if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
            n64 = O.StepNoise(eps=[e.to(dev).double() for e in noise.eps], perm=noise.perm.to(dev),
                              g_masks=[({k: v.to(dev).double() for k, v in m.items()} if m else None) for m in noise.g_masks], d_masks=[None, None, None])
            g64, d64 = f64(gsd), f64(dsd)
            want = O.train_iter_gan_oracle(cfg, 0, g64, d64, synth.zeros_like_opt(g64), synth.zeros_like_opt(d64), 1,
                                           inp['in_text'], inp['in_audio'].double(), inp['target'].double(), inp['vid'], n64)
            sd = G.state_dict()
            worst = 0.0
            tol_med = 2e-6 if mode == 'fp32' else 2e-4
            for k, gr in want['g_grads'].items():
                g = gr.detach().clone()
                dist.all_reduce(g)
                g /= world
                p, _, _ = O.adam_step(g64[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, cfg.learning_rate)
                d = (sd[k].double() - p).abs()
                ok = ok and d.max().item() <= 2.2 * cfg.learning_rate + 1e-6          # round-off gradients may flip a +-lr step
                if k not in ZERO_GRAD_KEYS:
                    worst = max(worst, d.median().item())
            res['median_' + mode] = worst
            res['ok_' + mode] = bool(ok and worst < tol_med)

        # (3) + (4): different initial weights per rank, graph-replayed steps with device-drawn noise
        config.set_mode('tf32'); config.set_graphs(True)
        torch.manual_seed(100 + rank)
        args, G, D, _, _ = build_ours(cfg, dev, seed=7 + rank)
        G.train(); D.train()
        g_opt = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
        d_opt = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
        ok = True
        for it in range(6):
            ret = TG.train_iter_gan(args, 11, inp['in_text'], inp['in_audio'], inp['target'], inp['vid'], G, D, g_opt, d_opt)
            ok = ok and all(torch.isfinite(torch.tensor(v)) for v in ret.values()) and same_across_ranks(G) and same_across_ranks(D)
        seeds = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(seeds, torch.tensor([G._noise.seed], dtype=torch.int64, device=dev))
        res['ok_graph'] = bool(ok)
        res['noise_seeds_distinct'] = len({int(s) for s in seeds}) == world
        res['graph_captured'] = any(s.graph is not None for s in G.engine()._gan_graph_slots.values())
        q.put((rank, res))
        q.close(); q.join_thread()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)          # no destroy_process_group() behind captured graphs that hold NCCL kernels: it does not return (round-2 call L)
    except BaseException as exc:
        import traceback
        q.put((rank, {'error': ''.join(traceback.format_exception(type(exc), exc, exc.__traceback__))[-2000:]}))
        q.close(); q.join_thread()
        os._exit(1)


@pytest.mark.gpu
def test_two_rank_nccl_train_iter_gan_matches_adam_on_mean_oracle_gradient():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    for rank, r in res:
        assert 'error' not in r, r['error']
        print('rank', rank, r)
        assert r['ok_fp32'] and r['ok_tf32'], (rank, r)
        assert r['ok_graph'] and r['noise_seeds_distinct'] and r['graph_captured'], (rank, r)
